// PTX wrappers and epilogue math shared by the tcgen05 kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace pyglm {

constexpr float kLoScale = 2048.0f;         // 2^11 between the two planes
constexpr float kRScale = 64.0f;            // residual planes carry r * 2^6
// exp nonlinearity, FP32 epilogue: e^x carries a relative error |x| 2^-24 and the residual planes saturate near
// lam = 10^6, so a column whose activation leaves x <= 16 (a rate of e^16 = 9e6 Hz) is flagged and the caller
// re-evaluates it on the FP64 path (pyglm_b200_ll_grad, PATH_AUTO); NaN / +inf activations are flagged too
constexpr float kExpSafe = 16.0f;
constexpr uint32_t kSw64 = 4;               // UMMA LayoutType::SWIZZLE_64B
constexpr uint32_t kSw128 = 2;              // UMMA LayoutType::SWIZZLE_128B

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): the hardware parks the thread until the phase completes or the time is up, so a
// waiting warp issues nothing in between (a bare poll loop competes with its scheduler's other warps for issue slots)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while; a poller must not)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
// (The watchdog counts polls instead of reading the clock: the 64-bit clock arithmetic of every poll was 11 % of all
// instructions the fused kernel's epilogue warps issued -- slots their schedulers' other warps needed.)
constexpr unsigned kMbarMaxPolls = 1u << 20;   // x up to 20 us per parked poll: a protocol bug traps within ~20 s
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned polls = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++polls > kMbarMaxPolls) __trap();
    }
}
// Wait of a single-thread role warp (MMA issuer): the warp scheduler favours the highest warp ids, and the role warps
// sit above the epilogue warps, so a polling role thread starves the lowest epilogue warps of its scheduler (measured:
// they reached the residual barrier ~2.2k cycles late).  A short nap between polls takes the thread out of arbitration.
#ifndef PYGLM_TC_ROLE_NAP
#define PYGLM_TC_ROLE_NAP 0
#endif
__device__ __forceinline__ void mbar_wait_role(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned polls = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (PYGLM_TC_ROLE_NAP) __nanosleep(PYGLM_TC_ROLE_NAP);
        if (++polls > kMbarMaxPolls) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in `mask` and each of those
// CTAs' barrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_multicast(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1,
                                                      uint16_t mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA spans two SMs, the leader CTA (cluster rank 0) issues it -----------
// shared::cluster address of `p` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// TMA load into this CTA's shared memory that signals a barrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
// same descriptor for a 128-byte-swizzled K-major operand: rows of 128 B, SBO = 1024 (8 rows), LBO unused
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)kSw128 << 61;
    return d;
}
// general form with an explicit layout type (kSw64 / kSw128)
__device__ __forceinline__ uint64_t umma_desc_layout(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)kSw64 << 61;
    return d;
}
// Wait used by warps whose waits are long (TMA producer): back off between polls so the spinning lane
// does not compete with the epilogue warps of its scheduler for issue / MIO slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns)
{
    if (mbar_try_wait(bar, parity)) return;
    unsigned polls = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (ns) __nanosleep(ns);
        if (++polls > kMbarMaxPolls) __trap();
    }
}

// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=F32 [4,6)=1, A/B=F16 (0),
// a_major [15], b_major [16] (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of FP32 accumulators -> 32 registers per thread (thread = TMEM lane)
// commit that arrives on the barrier at the same offset of every CTA in `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v)
{
    if constexpr (NC == 32) tmem_ld32(taddr, v);
    else if constexpr (NC == 16) tmem_ld16(taddr, v);
    else tmem_ld8(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Epilogue math (FP32): Poisson term and residual for one bin.
//   softplus: lam = log(1+e^x), f' = sigmoid(x)     (nlin.py:43)     exp: lam = f' = e^x (nlin.py:25)
// ---------------------------------------------------------------------------------------------
// Branches are decided per warp (votes), so the hot path has no divergence bookkeeping:
//   * softplus with every lane at x > 17.5: log(1+e^x) rounds to x and sigmoid(x) to 1 in FP32
//     (e^-17.5 < 2^-25), so no transcendental is evaluated at all -- the regime of a population
//     firing at tens of Hz (bias ~ 20, models/standard_glm.py:16-21);
//   * the log / reciprocal needed where a spike occurred are evaluated only if some lane has one.
template <int NLIN>
__device__ __forceinline__ void poisson_terms(float x, float s, float dt, float& term, float& r)
{
    const bool any_spike = __any_sync(0xffffffffu, s != 0.f);
    if constexpr (NLIN == PYGLM_B200_NLIN_EXP) {
        const float lam = expf(x);
        term = fmaf(-dt, lam, s * x);
        r = fmaf(-dt, lam, s);
    } else {
        float lam, sig;
        if (__all_sync(0xffffffffu, x > 17.5f)) {
            lam = x;
            sig = 1.0f;
        } else {
            const float e = __expf(-fabsf(x));                 // in [0,1]; flushes to 0 below x ~ -87
            // q = log1p(e)/e, so that for x < 0 lam = e q, log(lam) = x + log(q) and f'/lam = 1/((1+e) q) stay
            // finite and accurate however negative x is (log(lam) -> x, f'/lam -> 1), as in FP64
            float q = 1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e));       // series, |err| < e^4/5
            if (__any_sync(0xffffffffu, e >= 0.03125f)) {      // log(u)/(u-1) undoes the rounding of u = 1+e
                const float u = 1.0f + e;
                const float big = __fdividef(__logf(u), u - 1.0f);
                q = e >= 0.03125f ? big : q;
            }
            const float l1p = e * q;
            lam = x > 0.f ? x + l1p : l1p;
            float inv1pe;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1pe) : "f"(1.0f + e));   // 1 ulp; f' enters r linearly
            sig = x > 0.f ? inv1pe : e * inv1pe;
            term = -dt * lam;
            r = -dt * sig;
            if (any_spike) {                                   // ~2% of bins have s != 0
                const float loglam = x > 0.f ? __logf(lam) : x + __logf(q);
                const float ratio = x > 0.f ? __fdividef(sig, lam) : __fdividef(inv1pe, q);
                term = fmaf(s, s != 0.f ? loglam : 0.f, term);
                r = fmaf(s, ratio, r);
            }
            return;
        }
        term = -dt * lam;
        r = -dt * sig;
        if (any_spike) {
            term = fmaf(s, s != 0.f ? __logf(lam) : 0.f, term);
            r = fmaf(__fdividef(s, lam), sig, r);
        }
    }
}

// butterfly transpose-reduce over the 32 lanes of a warp for NC per-thread values:
// lane l ends with sum over lanes of v[l % NC]
template <int NC>
__device__ __forceinline__ float warp_column_sums(float* v, int lane)
{
#pragma unroll
    for (int off = 16; off >= NC; off >>= 1) {           // more lanes than columns: plain all-reduce steps
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
#pragma unroll
    for (int off = (NC < 32 ? NC / 2 : 16); off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float mine = up ? v[i + off] : v[i];
            const float other = up ? v[i] : v[i + off];
            v[i] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
    }
    return v[0];
}


// power of two s with  max*s in [2^14, 2^15)
__device__ __forceinline__ float pow2_scale(float mx)
{
    if (!(mx > 0.f) || !isfinite(mx)) return 1.0f;
    int e;
    frexpf(mx, &e);                                   // mx = m * 2^e, m in [0.5, 1)
    e = 15 - e;
    e = max(-100, min(100, e));
    return ldexpf(1.0f, e);
}

}  // namespace pyglm
