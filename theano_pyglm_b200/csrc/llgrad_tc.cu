// K2 (tensor-core path): fused population ll + gradient on tcgen05, one pass over X.
//
// Algebra (same as llgrad_simt.cu; glm.py:31-52, impulse.py:58, T.grad of glm.ll):
//   act[t][n] = bias[n] + sum_j X[t][j] M[j][n]          forward contraction   (K = N*B features)
//   r[t][n]   = (S/lam - dt) f'(act)                      epilogue
//   G[j][n]   = sum_t X[t][j] r[t][n]                     gradient contraction  (K = time)
//
// Both contractions read the same X tile, so one persistent CTA per SM streams X exactly once:
// a 128-bin tile arrives by TMA, the forward MMA accumulates act in TMEM, four epilogue warps
// turn it into lam / Poisson terms / residuals and write the residual tile back to shared
// memory as the B operand of the gradient MMA, which reuses the X tile still resident in
// shared memory as its (MN-major) A operand.  X never makes a second trip from HBM.
//
// Precision: tensor cores multiply 11-bit significands.  Every operand is therefore carried as
// an error-free two-term split v*s = v1 + v2*2^-11 with v1, v2 FP16 and s a power of two
// (22 significant bits, the precision of a 3xTF32 split, at twice the MMA rate and half the
// shared memory).  The three significant products x1*m1, x2*m1, x1*m2 are accumulated in FP32
// TMEM -- the 2^-11-scaled pair in its own accumulator -- and per-tile results are folded into
// FP64 registers, so nothing is ever accumulated in FP32 across more than one tile.
// X is split once per dataset (planes live in HBM, 4 bytes/element like FP32); M and r are split
// per call / per tile.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "llgrad_tc.cuh"

namespace pyglm {

// ---------------------------------------------------------------------------------------------
// Tile geometry
// ---------------------------------------------------------------------------------------------
constexpr int kTileT = 128;                 // time bins per tile == UMMA M of the forward MMA
constexpr int kChunkF = 32;                 // features per shared-memory chunk (64 B rows, SWIZZLE_64B)
constexpr int kMaxChunks = 5;               // <= 160 features on this fused path
constexpr int kNcol = 32;                   // postsynaptic columns per launch == UMMA N
constexpr int kChunkBytes = kTileT * 64;    // 8192: one X-plane chunk [128 rows][32 halves]
constexpr int kMChunkBytes = kNcol * 64;    // 2048: one M-plane chunk [32 rows][32 halves]
constexpr int kRBytes = kTileT * 64;        // 8192: one residual plane [128 rows][32 halves]
constexpr int kStages = 2;
constexpr int kTmemCols = 256;
constexpr int kThreads = 192;               // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
constexpr float kLoScale = 2048.0f;         // 2^11 between the two planes
constexpr float kRScale = 64.0f;            // residual planes carry r * 2^6
constexpr uint32_t kSw64 = 4;               // UMMA LayoutType::SWIZZLE_64B

void TcWorkspace::release()
{
    cudaFree(X1); cudaFree(X2); cudaFree(sx); cudaFree(colmax); cudaFree(Mp); cudaFree(colpar); cudaFree(part);
    X1 = X2 = nullptr; sx = nullptr; colmax = nullptr; Mp = nullptr; colpar = nullptr; part = nullptr;
    part_elems = 0; planes_ready = false;
    if (tmaps) free(tmaps);
    tmaps = nullptr;
}

bool tc_supported(int64_t T, int N, int B, int x_dtype)
{
    return x_dtype == PYGLM_B200_X_F32 && T > 0 && (int64_t)N * B <= kMaxChunks * kChunkF;
}

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
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)kSw64 << 61;
    return d;
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// One-time: per-feature scales and the FP16 split planes of X
// ---------------------------------------------------------------------------------------------
__global__ void tc_colmax_kernel(const float* __restrict__ X, int64_t T, int NB, int64_t ldx, unsigned* colmax)
{
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= NB) return;
    float m = 0.f;
    for (int64_t t = blockIdx.x; t < T; t += gridDim.x) m = fmaxf(m, fabsf(X[t * ldx + j]));
    atomicMax(&colmax[j], __float_as_uint(m));       // non-negative floats order like unsigned ints
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

__global__ void tc_scales_kernel(const unsigned* colmax, int NB, float* sx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < NB) sx[j] = pow2_scale(__uint_as_float(colmax[j]));
}

__global__ void tc_split_X_kernel(const float* __restrict__ X, int64_t T, int NB, int64_t ldx,
                                  const float* __restrict__ sx, __half* __restrict__ X1, __half* __restrict__ X2,
                                  int64_t ldp)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * ldp) return;
    const int64_t t = idx / ldp;
    const int j = (int)(idx - t * ldp);
    __half h1 = __float2half_rn(0.f), h2 = h1;
    if (j < NB) {
        const float v = X[t * ldx + j] * sx[j];       // exact: power-of-two scale
        h1 = __float2half_rn(v);
        h2 = __float2half_rn((v - __half2float(h1)) * kLoScale);
    }
    X1[idx] = h1;
    X2[idx] = h2;
}

// ---------------------------------------------------------------------------------------------
// Per call: scaled weight matrix M'[j][n] = A W w[n][j] * sm[n] / sx[j] as K-major split planes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_prep_M_kernel(const double* __restrict__ w, const int8_t* __restrict__ A, const double* __restrict__ W,
                 const double* __restrict__ bias, const float* __restrict__ sx,
                 int N, int B, int n_lo, int ncols, int Kp, __half* __restrict__ Mp, float* __restrict__ colpar)
{
    __shared__ float smax[256];
    const int nl = blockIdx.x;                 // 0..31
    const int NB = N * B;
    const bool live = nl < ncols;
    const int n = n_lo + nl;
    float mx = 0.f;
    if (live) {
        for (int j = threadIdx.x; j < NB; j += 256) {
            const int pre = j / B;
            const double a = A ? (double)A[(int64_t)pre * N + n] : 1.0;
            const double ww = W ? W[(int64_t)pre * N + n] : 1.0;
            const double m = (a * ww) * w[(int64_t)n * NB + j] / (double)sx[j];
            mx = fmaxf(mx, fabsf((float)m));
        }
    }
    smax[threadIdx.x] = mx;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + off]);
        __syncthreads();
    }
    const float sm = pow2_scale(smax[0] * 1.0001f);
    __half* M1 = Mp + (int64_t)nl * Kp;
    __half* M2 = Mp + (int64_t)(kNcol + nl) * Kp;
    for (int j = threadIdx.x; j < Kp; j += 256) {
        __half h1 = __float2half_rn(0.f), h2 = h1;
        if (live && j < NB) {
            const int pre = j / B;
            const double a = A ? (double)A[(int64_t)pre * N + n] : 1.0;
            const double ww = W ? W[(int64_t)pre * N + n] : 1.0;
            const double v = (a * ww) * w[(int64_t)n * NB + j] / (double)sx[j] * (double)sm;
            h1 = __double2half(v);
            h2 = __double2half((v - (double)__half2float(h1)) * (double)kLoScale);
        }
        M1[j] = h1;
        M2[j] = h2;
    }
    if (threadIdx.x == 0) {
        colpar[nl] = live ? 1.0f / sm : 0.f;
        colpar[kNcol + nl] = live ? (float)bias[n] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// Epilogue math (FP32): Poisson term and residual for one bin.
//   softplus: lam = log(1+e^x), f' = sigmoid(x)     (nlin.py:43)     exp: lam = f' = e^x (nlin.py:25)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void poisson_terms(float x, float s, float dt, int nlin, float& term, float& r)
{
    if (nlin == PYGLM_B200_NLIN_EXP) {
        const float lam = expf(x);
        term = fmaf(-dt, lam, s * x);
        r = fmaf(-dt, lam, s);
        return;
    }
    const float ax = fabsf(x);
    const float e = __expf(-ax);                       // in (0,1]
    float l1p;
    if (e < 0.03125f) l1p = e * (1.0f - e * (0.5f - e * (0.33333334f - 0.25f * e)));   // log1p series, |err| < e^5/5
    else              l1p = log1pf(e);
    const float lam = x > 0.f ? x + l1p : l1p;
    const float inv1pe = __frcp_rn(1.0f + e);
    const float sig = x > 0.f ? inv1pe : e * inv1pe;
    term = -dt * lam;
    r = -dt * sig;
    if (s != 0.f) {
        term = fmaf(s, logf(lam), term);
        r = fmaf(__fdiv_rn(s, lam), sig, r);
    }
}

// butterfly transpose-reduce: lane n ends with sum over the warp's 32 lanes of v[n]
__device__ __forceinline__ float warp_column_sums(float* v, int lane)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
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

struct TcKernelArgs {
    const uint8_t* S; int64_t T; int N; int halo;
    float dt; int nlin; int n_lo, ncols;
    int nch;                       // 32-feature chunks (1..5)
    int nmt;                       // 128-feature gradient tiles (1..2)
    int64_t ntiles;
    const __half* Mp; int Kp;      // [2][32][Kp]
    const float* colpar;           // [2][32]
    double* part;                  // per CTA: [nmt][128][32] G, then [32] ll, [32] g_bias
};

struct TcSmem {
    int stage_bytes;               // 2 * nch * kChunkBytes
    int off_M, off_R, off_bar, total;
};

__host__ __device__ inline TcSmem tc_smem_layout(int nch)
{
    TcSmem L;
    L.stage_bytes = 2 * nch * kChunkBytes;
    L.off_M = kStages * L.stage_bytes;
    L.off_R = L.off_M + 2 * nch * kMChunkBytes;
    L.off_bar = L.off_R + 2 * kRBytes;
    // slack so the second gradient tile's descriptor (4 chunks from chunk 4) stays inside the allocation
    int end = L.off_bar + 256;
    const int reach = (kStages - 1) * L.stage_bytes + nch * kChunkBytes + 8 * kChunkBytes;
    L.total = end > reach ? end : reach;
    return L;
}

__global__ void __launch_bounds__(kThreads, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2, TcKernelArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcSmem L = tc_smem_layout(a.nch);
    unsigned char* sM = smem + L.off_M;
    unsigned char* sR = smem + L.off_R;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* bar_full = bars;            // [2] TMA landed
    uint64_t* bar_empty = bars + 2;       // [2] stage consumed by both MMAs
    uint64_t* bar_fwd_full = bars + 4;    // activation accumulators ready
    uint64_t* bar_fwd_empty = bars + 5;   // ... drained by the epilogue
    uint64_t* bar_r_ready = bars + 6;     // residual planes written
    uint64_t* bar_bwd_full = bars + 7;    // gradient accumulators ready
    uint64_t* bar_bwd_empty = bars + 8;   // ... drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = a.nch;

    if (threadIdx.x == 0) {
        mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1);
        mbar_init(&bar_empty[0], 1); mbar_init(&bar_empty[1], 1);
        mbar_init(bar_fwd_full, 1); mbar_init(bar_fwd_empty, 4);
        mbar_init(bar_r_ready, 4);
        mbar_init(bar_bwd_full, 1); mbar_init(bar_bwd_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weight planes -> shared memory, K-major rows of 64 B with the 64B swizzle applied by hand
    {
        const int per_plane = nch * kMChunkBytes;            // bytes
        const int n16 = 2 * per_plane / 16;                  // 16-byte units in both planes
        for (int u = threadIdx.x; u < n16; u += kThreads) {
            const int plane = u / (per_plane / 16);
            const int v = u - plane * (per_plane / 16);
            const int c = v / (kMChunkBytes / 16);           // chunk
            const int w = v - c * (kMChunkBytes / 16);
            const int row = w >> 2, q = w & 3;               // row n (64 B), 16-byte unit within the row
            const uint4 val = *reinterpret_cast<const uint4*>(
                a.Mp + ((int64_t)(plane * kNcol + row) * a.Kp + c * kChunkF + q * 8));
            *reinterpret_cast<uint4*>(sM + plane * per_plane + c * kMChunkBytes + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = val;
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int it = 0;
            for (int64_t tile = first; tile < a.ntiles; tile += step, ++it) {
                const int s = it & 1;
                mbar_wait(&bar_empty[s], ((it >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_full[s], (uint32_t)L.stage_bytes);
                unsigned char* st = smem + s * L.stage_bytes;
                const int row0 = (int)(tile * kTileT);
                for (int c = 0; c < nch; ++c) {
                    tma_load_2d(st + c * kChunkBytes, &tmap1, &bar_full[s], c * kChunkF, row0);
                    tma_load_2d(st + (nch + c) * kChunkBytes, &tmap2, &bar_full[s], c * kChunkF, row0);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            constexpr uint32_t idesc_fwd = umma_idesc(128, kNcol, 0, 0);     // A: X K-major,  B: M K-major
            constexpr uint32_t idesc_bwd = umma_idesc(128, kNcol, 1, 1);     // A: X MN-major, B: r MN-major
            const uint32_t sM1 = smem_u32(sM), sM2 = sM1 + nch * kMChunkBytes;
            const uint32_t sR1 = smem_u32(sR), sR2 = sR1 + kRBytes;
            const uint32_t t_da = tmem_base, t_db = tmem_base + 32;
            int it = 0;
            for (int64_t tile = first; tile < a.ntiles; tile += step, ++it) {
                const int s = it & 1;
                const uint32_t ph = it & 1;
                const uint32_t sX1 = smem_u32(smem + s * L.stage_bytes), sX2 = sX1 + nch * kChunkBytes;
                mbar_wait(&bar_full[s], (it >> 1) & 1);
                mbar_wait(bar_fwd_empty, ph ^ 1);
                tc_fence_after();
                // ---- forward: act = X1 M1 (t_da) ; X2 M1 + X1 M2 (t_db, carries 2^-11)
                for (int c = 0; c < nch; ++c) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t acc = (c | ks) ? 1u : 0u;
                        const uint64_t dx1 = umma_desc(sX1 + c * kChunkBytes + ks * 32, 16, 512);
                        const uint64_t dx2 = umma_desc(sX2 + c * kChunkBytes + ks * 32, 16, 512);
                        const uint64_t dm1 = umma_desc(sM1 + c * kMChunkBytes + ks * 32, 16, 512);
                        const uint64_t dm2 = umma_desc(sM2 + c * kMChunkBytes + ks * 32, 16, 512);
                        umma_f16(t_da, dx1, dm1, idesc_fwd, acc);
                        umma_f16(t_db, dx2, dm1, idesc_fwd, acc);
                        umma_f16(t_db, dx1, dm2, idesc_fwd, 1u);
                    }
                }
                umma_commit(bar_fwd_full);
                // ---- gradient: G = X1^T r1 (ga) ; X2^T r1 + X1^T r2 (gb), X tile reused from smem
                mbar_wait(bar_r_ready, ph);
                mbar_wait(bar_bwd_empty, ph ^ 1);
                tc_fence_after();
                for (int mt = 0; mt < a.nmt; ++mt) {
                    const uint32_t t_ga = tmem_base + 64 + mt * 64, t_gb = t_ga + 32;
#pragma unroll
                    for (int ks = 0; ks < kTileT / 16; ++ks) {
                        const uint32_t acc = ks ? 1u : 0u;
                        const uint32_t xo = mt * 4 * kChunkBytes + ks * 1024;    // 4 chunks per tile, 16 bins per step
                        const uint64_t dx1 = umma_desc(sX1 + xo, kChunkBytes, 512);
                        const uint64_t dx2 = umma_desc(sX2 + xo, kChunkBytes, 512);
                        const uint64_t dr1 = umma_desc(sR1 + ks * 1024, kRBytes, 512);
                        const uint64_t dr2 = umma_desc(sR2 + ks * 1024, kRBytes, 512);
                        umma_f16(t_ga, dx1, dr1, idesc_bwd, acc);
                        umma_f16(t_gb, dx2, dr1, idesc_bwd, acc);
                        umma_f16(t_gb, dx1, dr2, idesc_bwd, 1u);
                    }
                }
                umma_commit(bar_bwd_full);
                umma_commit(&bar_empty[s]);
            }
        }
    } else {
        // ================================ epilogue warps ==============================
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                   // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float inv_sm = a.colpar[lane], bias_l = a.colpar[kNcol + lane];   // column `lane`
        double gacc[2][kNcol];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int c = 0; c < kNcol; ++c) gacc[mt][c] = 0.0;
        double ll_acc = 0.0, gb_acc = 0.0;               // lane n accumulates column n

        int it = 0;
        for (int64_t tile = first; tile < a.ntiles; tile += step, ++it) {
            const uint32_t ph = it & 1;
            const int64_t t = tile * kTileT + row;
            const bool live = t < a.T;
            // spikes of this bin for the requested columns (27 contiguous bytes at C2)
            uint32_t sp[kNcol / 4];
#pragma unroll
            for (int i = 0; i < kNcol / 4; ++i) sp[i] = 0;
            if (live) {
                const uint8_t* srow = a.S + ((int64_t)a.halo + t) * a.N + a.n_lo;
#pragma unroll
                for (int c = 0; c < kNcol; ++c)
                    if (c < a.ncols) sp[c >> 2] |= (uint32_t)srow[c] << ((c & 3) * 8);
            }
            mbar_wait(bar_fwd_full, ph);
            tc_fence_after();
            float da[kNcol], db[kNcol];
            tmem_ld32(t_lane + 0, da);
            tmem_ld32(t_lane + 32, db);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_fwd_empty);

            // act, Poisson term, residual; residual planes back to smem (MN-major, 64B swizzle)
            uint32_t r1p[kNcol / 2], r2p[kNcol / 2];
#pragma unroll
            for (int c = 0; c < kNcol; ++c) {
                const float ism = __shfl_sync(0xffffffffu, inv_sm, c);
                const float bc = __shfl_sync(0xffffffffu, bias_l, c);
                float term = 0.f, r = 0.f;
                if (live && c < a.ncols) {
                    const float x = fmaf(fmaf(db[c], 1.0f / kLoScale, da[c]), ism, bc);
                    const float s = (float)((sp[c >> 2] >> ((c & 3) * 8)) & 0xffu);
                    poisson_terms(x, s, a.dt, a.nlin, term, r);
                }
                da[c] = term;                            // reuse registers: da <- ll terms, db <- residuals
                db[c] = r;
                const float rs = r * kRScale;
                const __half h1 = __float2half_rn(rs);
                const __half h2 = __float2half_rn((rs - __half2float(h1)) * kLoScale);
                const uint32_t u1 = __half_as_ushort(h1), u2 = __half_as_ushort(h2);
                if (c & 1) { r1p[c >> 1] |= u1 << 16; r2p[c >> 1] |= u2 << 16; }
                else       { r1p[c >> 1] = u1;        r2p[c >> 1] = u2; }
            }
            {
                unsigned char* d1 = sR + row * 64;
                unsigned char* d2 = sR + kRBytes + row * 64;
                const int sw = (row >> 1) & 3;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    *reinterpret_cast<uint4*>(d1 + ((u ^ sw) << 4)) = make_uint4(r1p[4 * u], r1p[4 * u + 1], r1p[4 * u + 2], r1p[4 * u + 3]);
                    *reinterpret_cast<uint4*>(d2 + ((u ^ sw) << 4)) = make_uint4(r2p[4 * u], r2p[4 * u + 1], r2p[4 * u + 2], r2p[4 * u + 3]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_r_ready);

            // column sums of the ll terms and residuals (bias gradient) while the gradient MMA runs
            ll_acc += (double)warp_column_sums(da, lane);
            gb_acc += (double)warp_column_sums(db, lane);

            // fold this tile's gradient block into FP64
            mbar_wait(bar_bwd_full, ph);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (mt < a.nmt) {
                    tmem_ld32(t_lane + 64 + mt * 64, da);
                    tmem_ld32(t_lane + 64 + mt * 64 + 32, db);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < kNcol; ++c)
                        gacc[mt][c] += (double)da[c] + (double)db[c] * (1.0 / kLoScale);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_bwd_empty);
        }

        // ---- per-CTA partials: G rows (feature = mt*128 + row), then ll / g_bias
        double* gp = a.part + (int64_t)blockIdx.x * ((int64_t)a.nmt * 128 * kNcol + 2 * kNcol);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            if (mt < a.nmt) {
#pragma unroll
                for (int c = 0; c < kNcol; ++c) gp[((int64_t)mt * 128 + row) * kNcol + c] = gacc[mt][c];
            }
        }
        double* sred = reinterpret_cast<double*>(sR);    // residual planes are idle now: [4][2][32]
        asm volatile("bar.sync 1, 128;" ::: "memory");   // all epilogue warps past their last MMA wait
        sred[(q * 2 + 0) * kNcol + lane] = ll_acc;
        sred[(q * 2 + 1) * kNcol + lane] = gb_acc;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q == 0) {
            double l = 0.0, g = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { l += sred[(k * 2 + 0) * kNcol + lane]; g += sred[(k * 2 + 1) * kNcol + lane]; }
            double* lp = gp + (int64_t)a.nmt * 128 * kNcol;
            lp[lane] = l;
            lp[kNcol + lane] = g;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
}

// Sum the per-CTA partials in CTA order and undo the scales.
__global__ void __launch_bounds__(256)
tc_final_kernel(const double* __restrict__ part, int nctas, int nmt, int N, int B, int n_lo, int ncols,
                const float* __restrict__ sx, const int8_t* __restrict__ A, const double* __restrict__ W,
                double* __restrict__ out_ll, double* __restrict__ out_gb, double* __restrict__ out_gw)
{
    const int64_t NB = (int64_t)N * B;
    const int64_t per_cta = (int64_t)nmt * 128 * kNcol + 2 * kNcol;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_gw && idx < NB * ncols) {
        const int nl = (int)(idx / NB);
        const int64_t j = idx - (int64_t)nl * NB;
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += part[c * per_cta + j * kNcol + nl];
        const int n = n_lo + nl, pre = (int)(j / B);
        const double a = A ? (double)A[(int64_t)pre * N + n] : 1.0;
        const double ww = W ? W[(int64_t)pre * N + n] : 1.0;
        out_gw[idx] = (a * ww) * s / ((double)sx[j] * (double)kRScale);
    }
    if (idx < 2 * ncols) {
        const int which = (int)(idx / ncols), nl = (int)(idx - (int64_t)which * ncols);
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += part[c * per_cta + (int64_t)nmt * 128 * kNcol + which * kNcol + nl];
        if (which == 0) out_ll[nl] = s;
        else if (out_gb) out_gb[nl] = s;               // residual column sums are carried unscaled
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_plane_map(CUtensorMap* map, const __half* base, int64_t T, int64_t ldp)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        PYGLM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return PYGLM_B200_ECUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ldp, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)ldp * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)kChunkF, (cuuint32_t)kTileT};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for T=%lld ldp=%lld", (int)r, (long long)T, (long long)ldp);
        return PYGLM_B200_ECUDA;
    }
    return PYGLM_B200_OK;
}

static int ensure_planes(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (ws.planes_ready) return PYGLM_B200_OK;
    const int NB = a.N * a.B;
    ws.ldp = round_up(NB, 8);
    int dev = 0;
    PYGLM_CUDA(cudaGetDevice(&dev));
    PYGLM_CUDA(cudaDeviceGetAttribute(&ws.num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t plane = (size_t)a.T * ws.ldp;
    PYGLM_CUDA(cudaMalloc(&ws.X1, plane * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.X2, plane * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.sx, NB * sizeof(float)));
    PYGLM_CUDA(cudaMalloc(&ws.colmax, NB * sizeof(unsigned)));
    PYGLM_CUDA(cudaMalloc(&ws.Mp, (size_t)2 * kNcol * kMaxChunks * kChunkF * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.colpar, 2 * kNcol * sizeof(float)));
    PYGLM_CUDA(cudaMemsetAsync(ws.colmax, 0, NB * sizeof(unsigned), stream));
    dim3 gmax((unsigned)std::min<int64_t>(a.T, 148 * 16), (unsigned)ceil_div(NB, 128));
    tc_colmax_kernel<<<gmax, 128, 0, stream>>>(a.X, a.T, NB, a.ldx, ws.colmax);
    PYGLM_CUDA(cudaGetLastError());
    tc_scales_kernel<<<(unsigned)ceil_div(NB, 128), 128, 0, stream>>>(ws.colmax, NB, ws.sx);
    PYGLM_CUDA(cudaGetLastError());
    tc_split_X_kernel<<<(unsigned)ceil_div((int64_t)plane, 256), 256, 0, stream>>>(a.X, a.T, NB, a.ldx, ws.sx, ws.X1, ws.X2, ws.ldp);
    PYGLM_CUDA(cudaGetLastError());
    ws.tmaps = malloc(2 * sizeof(CUtensorMap));
    if (!ws.tmaps) { set_error("host allocation failed"); return PYGLM_B200_ENOMEM; }
    CUtensorMap* maps = static_cast<CUtensorMap*>(ws.tmaps);
    int rc;
    if ((rc = make_plane_map(&maps[0], ws.X1, a.T, ws.ldp))) return rc;
    if ((rc = make_plane_map(&maps[1], ws.X2, a.T, ws.ldp))) return rc;
    ws.planes_ready = true;
    return PYGLM_B200_OK;
}

int launch_tc_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (a.T <= 0 || a.ncols <= 0) return PYGLM_B200_OK;
    int rc = ensure_planes(a, ws, stream);
    if (rc) return rc;
    const int NB = a.N * a.B;
    const int nch = (int)ceil_div(NB, kChunkF);
    const int nmt = (int)ceil_div(nch, 4);
    const int Kp = nch * kChunkF;
    const int64_t ntiles = ceil_div(a.T, kTileT);
    const int nctas = (int)std::min<int64_t>(ntiles, ws.num_sms);
    const size_t per_cta = (size_t)nmt * 128 * kNcol + 2 * kNcol;
    if (ws.part_elems < per_cta * nctas) {
        cudaFree(ws.part);
        ws.part = nullptr; ws.part_elems = 0;
        PYGLM_CUDA(cudaMalloc(&ws.part, per_cta * nctas * sizeof(double)));
        ws.part_elems = per_cta * nctas;
    }
    const TcSmem L = tc_smem_layout(nch);
    const int smem_bytes = L.total + 1024;
    static int smem_set = 0;
    if (smem_set < smem_bytes) {
        PYGLM_CUDA(cudaFuncSetAttribute(tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        smem_set = smem_bytes;
    }
    const CUtensorMap* maps = static_cast<const CUtensorMap*>(ws.tmaps);
    for (int c0 = 0; c0 < a.ncols; c0 += kNcol) {
        const int nc = std::min(kNcol, a.ncols - c0);
        const int n_lo = a.n_lo + c0;
        tc_prep_M_kernel<<<kNcol, 256, 0, stream>>>(a.w, a.A, a.W, a.bias, ws.sx, a.N, a.B, n_lo, nc, Kp, ws.Mp, ws.colpar);
        PYGLM_CUDA(cudaGetLastError());
        TcKernelArgs k{};
        k.S = a.S; k.T = a.T; k.N = a.N; k.halo = a.halo; k.dt = (float)a.dt; k.nlin = a.nlin;
        k.n_lo = n_lo; k.ncols = nc; k.nch = nch; k.nmt = nmt; k.ntiles = ntiles;
        k.Mp = ws.Mp; k.Kp = Kp; k.colpar = ws.colpar; k.part = ws.part;
        tc_fused_kernel<<<nctas, kThreads, smem_bytes, stream>>>(maps[0], maps[1], k);
        PYGLM_CUDA(cudaGetLastError());
        const int64_t work = std::max<int64_t>((int64_t)NB * nc, 2 * nc);
        tc_final_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, stream>>>(
            ws.part, nctas, nmt, a.N, a.B, n_lo, nc, ws.sx, a.A, a.W,
            a.out_ll + c0, a.out_gb ? a.out_gb + c0 : nullptr, a.out_gw ? a.out_gw + (int64_t)c0 * NB : nullptr);
        PYGLM_CUDA(cudaGetLastError());
    }
    return PYGLM_B200_OK;
}

}  // namespace pyglm
