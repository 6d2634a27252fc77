// K2 (tensor-core path, large populations): two tcgen05 GEMM kernels for N*B > 160 features.
//
// When the X tile of all features no longer fits in shared memory next to its pipeline (the fused
// kernel in llgrad_tc.cu), the two contractions become two kernels, each a classic TMA -> tcgen05 ->
// TMEM -> epilogue pipeline over the same FP16 split planes:
//
//   tc_gemm_fwd_kernel   act tile [128 bins x 128 neurons] = X[128 x K] M[K x 128], K = N*B streamed in
//                        32-feature chunks through a 6-stage ring; the epilogue (FP32) turns the TMEM
//                        accumulators into Poisson terms and residuals r, writes r to HBM as split FP16
//                        planes and keeps per-column ll / g_bias sums in FP64.          glm.py:39-52
//   tc_gemm_bwd_kernel   G tile [128 features x 128 neurons] = sum_t X[t]^T r[t] over this CTA's share of
//                        the recording (split-K over time); both operands MN-major straight from TMA;
//                        every 128 bins the TMEM block is folded into FP64 registers (TMEM FP32
//                        accumulation is not clean over long sums, see llgrad_tc.cu).   T.grad(glm.ll)
//
// Same operand splitting as the fused kernel: x1 m1 in one accumulator, x2 m1 + x1 m2 (carrying 2^-11) in
// a second one.  A 128x128x16 FP16 MMA reads 8 KB of operands for 64 cycles of tensor pipe, so these tiles
// are tensor-bound rather than shared-memory-bound.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "llgrad_tc.cuh"
#include "tc_common.cuh"

#ifndef PYGLM_GEMM_EXPERIMENTS
#define PYGLM_GEMM_EXPERIMENTS 0
#endif

namespace pyglm {

constexpr int kGT = 128;                     // bins per tile (UMMA M of the forward, K block of the gradient)
constexpr int kGN = 128;                     // neurons per tile (UMMA N)
constexpr int kGF = 128;                     // features per gradient tile (UMMA M of the gradient)
constexpr int kGEpiWarps = 16;
// Epilogue warps come first and the two single-thread role warps last: the warp scheduler favours the highest
// warp id, and the TMA / MMA issuers must never wait behind sixteen epilogue warps for an issue slot.
constexpr int kGTmaWarp = kGEpiWarps;
constexpr int kGMmaWarp = kGEpiWarps + 1;
constexpr int kGThreads = 32 * (kGEpiWarps + 2);
constexpr int kGColsPerWarp = kGN / 4;       // 32 columns per epilogue warp (4 column groups x 4 lane quarters)
constexpr int kFwdStages = 7;                // 7 x 32 KB + barriers = 225 KB of the 227 KB a CTA may use
constexpr int kFwdSegChunks = 8;             // forward: 256 features per TMEM accumulation segment ...
constexpr int kFwdSingleSegmentChunks = 48;  // ... when there are more than 1536 features
constexpr int kFwdStageBytes = 4 * kGT * 64; // X1, X2, M1, M2 chunks of [128 rows][32 halves]
constexpr int kBwdStages = 3;
constexpr int kBwdRows = 64;                 // bins per gradient stage
constexpr int kBwdChunkBytes = kBwdRows * 64;
constexpr int kBwdStageBytes = 16 * kBwdChunkBytes;   // X1, X2, r1, r2: 4 chunks each

struct GemmWorkspace {
    __half* Mp = nullptr; size_t Mp_elems = 0;         // [2][Npr][Kp]
    float* colpar = nullptr; size_t colpar_elems = 0;  // [2][Npr]
    __half* R = nullptr; size_t R_elems = 0;           // [2][T][Npr]
    double* part = nullptr; size_t part_elems = 0;     // forward: [nctas][4][Npr][2]
    double* Gp = nullptr; size_t Gp_elems = 0;         // gradient: [splits][NBp][Npr]
    CUtensorMap mapX64[2];                             // X planes with 64-row boxes (gradient kernel)
    bool mapX64_ready = false;
};

static int ensure(void** p, size_t* have, size_t want, size_t esz)
{
    if (*have >= want) return PYGLM_B200_OK;
    cudaFree(*p);
    *p = nullptr; *have = 0;
    note_allocation();
    PYGLM_CUDA(cudaMalloc(p, want * esz));
    *have = want;
    return PYGLM_B200_OK;
}

void tc_gemm_release(void* w)
{
    GemmWorkspace* g = static_cast<GemmWorkspace*>(w);
    if (!g) return;
    cudaFree(g->Mp); cudaFree(g->colpar); cudaFree(g->R); cudaFree(g->part); cudaFree(g->Gp);
    delete g;
}

// ---------------------------------------------------------------------------------------------
// scaled weight planes for all requested columns: block = one column
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_gemm_prep_M_kernel(const double* __restrict__ w, const int8_t* __restrict__ A, const double* __restrict__ W,
                      const double* __restrict__ bias, const float* __restrict__ sx,
                      int N, int B, int F, int n_lo, int ncols, int Npr, int Kp, __half* __restrict__ Mp, float* __restrict__ colpar)
{
    __shared__ float smax[256];
    const int nl = blockIdx.x;
    const int NS = N * B, NB = NS + F;         // spike-history features, all features
    const bool live = nl < ncols;
    const int n = n_lo + nl;
    float mx = 0.f;
    if (live) {
        for (int j = threadIdx.x; j < NB; j += 256) {
            const int pre = j / B;
            const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
            const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
            mx = fmaxf(mx, fabsf((float)((a * ww) * w[(int64_t)n * NB + j] / (double)sx[j])));
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
    __half* M2 = Mp + ((int64_t)Npr + nl) * Kp;
    for (int j = threadIdx.x; j < Kp; j += 256) {
        __half h1 = __float2half_rn(0.f), h2 = h1;
        if (live && j < NB) {
            const int pre = j / B;
            const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
            const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
            const double v = (a * ww) * w[(int64_t)n * NB + j] / (double)sx[j] * (double)sm;
            h1 = __double2half(v);
            h2 = __double2half((v - (double)__half2float(h1)) * (double)kLoScale);
        }
        M1[j] = h1;
        M2[j] = h2;
    }
    if (threadIdx.x == 0) {
        colpar[nl] = live ? 1.0f / sm : 0.f;
        colpar[Npr + nl] = live ? (float)bias[n] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// forward GEMM + epilogue
// ---------------------------------------------------------------------------------------------
struct GemmFwdArgs {
    const uint8_t* Sp; int Nps;        // padded spikes [T][Nps]
    int64_t T; int n_lo, ncols, Npr;
    int nkc;                           // 32-feature chunks
    int seg;                           // chunks per TMEM accumulation segment (see the MMA issuer)
    int debug;                         // experiments (PYGLM_GEMM_DEBUG): 1 no MMAs, 2 always the same tile (L2-hot loads), 4 no epilogue math
    int64_t ntt; int ncb;              // time tiles, column blocks
    const float* colpar;               // [2][Npr]
    __half* R; int64_t plane;          // residual planes: R1 at R, R2 at R + plane; row pitch Npr
    double* part;                      // [nctas][4 quarters][Npr][2]
    float dt;
    unsigned* flags;                   // [N] range flags (exp nonlinearity, kExpSafe) or nullptr
};

// MODE 0: one CTA per 128 x 128 tile.
// MODE 1 (experiment, PYGLM_GEMM_PAIR=1): clusters of two CTAs on the column blocks 2j, 2j+1 of one time tile; the X
//         chunk is the same for both, so each CTA fetches half of its rows and multicasts them into both stages
//         (X maps with 64-row boxes; a stage is refilled when BOTH consumers are done: commits go to both CTAs).
// MODE 2: CTA pair (cta_group::2).  The pair owns 256 bins x 128 neurons: each CTA holds its own 128 bins of X and
//         only 64 of the 128 weight columns, the leader issues 256 x 128 x 16 MMAs that read both halves, and each
//         CTA's TMEM receives the accumulators of its own bins.  A CTA ingests 24 KB per K chunk instead of 32 KB
//         for the same outputs: the forward kernel is bound by what an SM can ingest.  All TMA transactions signal
//         the leader's full barrier; the leader's commits release the stages and publish the accumulators in both
//         CTAs; both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.  (M maps: 64-row boxes.)
constexpr int kFwdStages2 = 9;
constexpr int kFwdStages3 = 3;               // MODE 3: 64-feature chunks with 128-byte rows (SWIZZLE_128B)
constexpr int kFwdStageBytes3 = 4 * kGT * 128;
constexpr int kFwdStageBytes2 = 2 * kGT * 64 + 2 * (kGN / 2) * 64;       // X1, X2 [128 rows], M1, M2 halves [64 rows]
static_assert(kFwdStages2 * kFwdStageBytes2 <= kFwdStages * kFwdStageBytes, "the pair layout fits in the same shared memory");

template <int NLIN, int MODE>
__global__ void __launch_bounds__(kGThreads, 1)
tc_gemm_fwd_kernel(const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
                   const __grid_constant__ CUtensorMap mapM1, const __grid_constant__ CUtensorMap mapM2, GemmFwdArgs a)
{
    constexpr bool PAIR = MODE == 1;
    constexpr bool DUO = MODE == 2;
    constexpr bool WIDE = MODE == 3;             // one CTA per tile like MODE 0, K chunks of 64 features in 128-byte rows
    constexpr bool CLUSTER = PAIR || DUO;
    constexpr int kStages = DUO ? kFwdStages2 : WIDE ? kFwdStages3 : kFwdStages;
    constexpr int kStageBytes = DUO ? kFwdStageBytes2 : WIDE ? kFwdStageBytes3 : kFwdStageBytes;
    constexpr int kRowB = WIDE ? 128 : 64;       // bytes per operand row in a stage
    constexpr int kChunkFeat = WIDE ? 64 : 32;
    const uint32_t crank = CLUSTER ? cluster_ctarank() : 0u;
    // work items of this CTA: w = first, first + stride, ...; item -> (time tile, column block)
    const int64_t w_first = CLUSTER ? blockIdx.x / 2 : blockIdx.x;
    const int64_t w_stride = CLUSTER ? gridDim.x / 2 : gridDim.x;
    const int cbw = PAIR ? a.ncb / 2 : a.ncb;                    // column-block work items per time tile (pair)
    const int64_t nwork = (DUO ? (a.ntt + 1) / 2 : a.ntt) * cbw;
    auto tile_tt = [&](int64_t w) { return DUO ? 2 * (w / cbw) + crank : w / cbw; };
    auto tile_cb = [&](int64_t w) { return PAIR ? 2 * (int)(w % cbw) + (int)crank : (int)(w % cbw); };
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* bar_full = bars;                       // [kStages]
    uint64_t* bar_empty = bars + kStages;            // [kStages]
    uint64_t* bar_acc_full = bars + 2 * kStages;     // [2]
    uint64_t* bar_acc_empty = bar_acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], PAIR ? 2 : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], DUO ? 2 * kGEpiWarps : kGEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (DUO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) cluster_sync_all();         // the peer's barriers and TMEM exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kGTmaWarp) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t s = 0, ph = 1;              // stage and the parity its empty barrier is waited on
            const uint32_t lead_full0 = DUO ? mapa_u32(&bar_full[0], 0) : 0u;
            for (int64_t w = w_first; w < nwork; w += w_stride) {
                const int row0 = (a.debug & 2) ? 0 : (int)(tile_tt(w) * kGT);
                const int col0 = tile_cb(w) * kGN;
                for (int kc = 0; kc < a.nkc; ++kc) {
                    mbar_wait(&bar_empty[s], ph);
                    unsigned char* st = smem + s * kStageBytes;
                    if constexpr (DUO) {         // every transaction of the pair lands on the leader's barrier
                        if (crank == 0) mbar_arrive_expect_tx(&bar_full[s], 2 * kStageBytes);
                        const uint32_t lead_full = lead_full0 + s * 8;
                        const int colh = col0 + (int)crank * (kGN / 2);
                        tma_load_2d_pair(st, &mapX1, lead_full, kc * 32, row0);
                        tma_load_2d_pair(st + kGT * 64, &mapX2, lead_full, kc * 32, row0);
                        tma_load_2d_pair(st + 2 * kGT * 64, &mapM1, lead_full, kc * 32, colh);
                        tma_load_2d_pair(st + 2 * kGT * 64 + (kGN / 2) * 64, &mapM2, lead_full, kc * 32, colh);
                    } else {
                        mbar_arrive_expect_tx(&bar_full[s], kStageBytes);
                        if (PAIR) {              // this CTA's half of the rows, into both CTAs
                            const int half = (int)crank * (kGT / 2);
                            tma_load_2d_multicast(st + half * 64, &mapX1, &bar_full[s], kc * 32, row0 + half, (uint16_t)3);
                            tma_load_2d_multicast(st + kGT * 64 + half * 64, &mapX2, &bar_full[s], kc * 32, row0 + half, (uint16_t)3);
                        } else {
                            tma_load_2d(st, &mapX1, &bar_full[s], kc * kChunkFeat, row0);
                            tma_load_2d(st + kGT * kRowB, &mapX2, &bar_full[s], kc * kChunkFeat, row0);
                        }
                        tma_load_2d(st + 2 * kGT * kRowB, &mapM1, &bar_full[s], kc * kChunkFeat, col0);
                        tma_load_2d(st + 3 * kGT * kRowB, &mapM2, &bar_full[s], kc * kChunkFeat, col0);
                    }
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == kGMmaWarp) {
        // ================================ MMA issuer ==================================
        if (lane == 0 && (!DUO || crank == 0)) {
            // acc0 = x1 m1 and acc1 = x1 m2 + x2 m1 sit in adjacent TMEM columns and M1 | M2 are adjacent in the
            // stage, so x1 multiplies both in ONE N=256 MMA: x1 is read from shared memory once, not twice
            constexpr uint32_t idesc = umma_idesc(kGT, kGN, 0, 0);
            constexpr uint32_t idesc2 = umma_idesc(kGT, 2 * kGN, 0, 0);
            constexpr uint32_t idesc_duo = umma_idesc(2 * kGT, kGN, 0, 0);          // 256 bins (two CTAs) x 128 neurons
            // FP32 accumulation in TMEM loses ~1 ulp of the running sum per MMA step, and the loss does not
            // average out: over K = 10^4 features the activation error reached 2e-6 relative.  So the K loop
            // is cut into segments of `seg` chunks that alternate between the two TMEM buffers; the epilogue
            // warps drain every finished segment into FP32 registers (round-to-nearest adds) while the next
            // one accumulates.  Short running sums are small, and so are their rounding losses.
            // this one thread feeds the tensor pipe, a K chunk every ~400 cycles: the loop carries its stage / phase /
            // segment counters and descriptor bases instead of recomputing them with divisions
            uint32_t sgc = 0, s = 0, ph = 0;
            const uint64_t d_stage0 = WIDE ? umma_desc_sw128(smem_u32(smem)) : umma_desc(smem_u32(smem), 16, 512);
            for (int64_t w = w_first; w < nwork; w += w_stride) {
                int in_seg = 0;
                for (int kc = 0; kc < a.nkc; ++kc) {
                    const int ab = sgc & 1;
                    const uint32_t t_a = tmem_base + ab * 256, t_b = t_a + 128;
                    if (in_seg == 0) {
                        mbar_wait(&bar_acc_empty[ab], ((sgc >> 1) & 1) ^ 1);
                        tc_fence_after();
                    }
                    mbar_wait(&bar_full[s], ph);
                    tc_fence_after();
                    const uint64_t dx1 = d_stage0 + (uint64_t)(s * (kStageBytes >> 4));
                    const uint64_t dx2 = dx1 + ((kGT * kRowB) >> 4);
                    const uint64_t dm1 = dx1 + ((2 * kGT * kRowB) >> 4);            // MODE 0/1/3: rows 128..255 of this tile are M2
                    const bool seg_end = in_seg == a.seg - 1 || kc == a.nkc - 1;
                    if constexpr (DUO) {
                        const uint64_t dm2 = dm1 + (((kGN / 2) * 64) >> 4);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t acc = (in_seg | ks) ? 1u : 0u;
                            umma_f16_pair(t_a, dx1 + 2 * ks, dm1 + 2 * ks, idesc_duo, acc);   // X1 M1
                            umma_f16_pair(t_b, dx2 + 2 * ks, dm1 + 2 * ks, idesc_duo, acc);   // X2 M1
                            umma_f16_pair(t_b, dx1 + 2 * ks, dm2 + 2 * ks, idesc_duo, 1u);    // X1 M2
                        }
                        umma_commit_pair(&bar_empty[s], (uint16_t)3);
                        if (seg_end) umma_commit_pair(&bar_acc_full[ab], (uint16_t)3);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < kChunkFeat / 16; ++ks) {
                            if (a.debug & 1) break;
                            const uint32_t acc = (in_seg | ks) ? 1u : 0u;
                            umma_f16(t_a, dx1 + 2 * ks, dm1 + 2 * ks, idesc2, acc);    // X1 [M1 | M2]
                            umma_f16(t_b, dx2 + 2 * ks, dm1 + 2 * ks, idesc, 1u);      // X2 M1
                        }
                        if (PAIR) umma_commit_multicast(&bar_empty[s], (uint16_t)3);
                        else umma_commit(&bar_empty[s]);
                        if (seg_end) umma_commit(&bar_acc_full[ab]);
                    }
                    if (seg_end) { ++sgc; in_seg = 0; } else ++in_seg;
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ================================ epilogue warps ==============================
        const int q = warp & 3;
        const int cg = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + cg * kGColsPerWarp;
        double* my_part = a.part + ((int64_t)blockIdx.x * 4 + q) * a.Npr * 2;
        const int nseg = (a.nkc + a.seg - 1) / a.seg;
        uint32_t sgc = 0;
        for (int64_t w = w_first; w < nwork; w += w_stride) {
            const int64_t t = tile_tt(w) * kGT + row;
            const int col0 = tile_cb(w) * kGN + cg * kGColsPerWarp;                 // first of this warp's 32 columns
            const float lv = t < a.T ? 1.0f : 0.0f;
            // column parameters: lane l holds column col0 + l
            const float ism_l = a.colpar[col0 + lane], bias_l = a.colpar[a.Npr + col0 + lane];
            uint32_t sp[8];
            if (t < a.T && col0 < a.ncols) {     // padded column groups: nothing to read (Sp has N rounded up to 32, +32)
                const uint4* src = reinterpret_cast<const uint4*>(a.Sp + t * a.Nps + a.n_lo + col0);
                if (((a.n_lo + col0) & 15) == 0) {
                    const uint4 v0 = src[0], v1 = src[1];
                    sp[0] = v0.x; sp[1] = v0.y; sp[2] = v0.z; sp[3] = v0.w; sp[4] = v1.x; sp[5] = v1.y; sp[6] = v1.z; sp[7] = v1.w;
                } else {
                    const uint8_t* sb = a.Sp + t * a.Nps + a.n_lo + col0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) sp[i] = sb[4 * i] | (sb[4 * i + 1] << 8) | (sb[4 * i + 2] << 16) | ((uint32_t)sb[4 * i + 3] << 24);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) sp[i] = 0;
            }
            // drain the segments: act[c] = sum over segments of (acc0 + acc1 2^-11), this lane's bin x 32 columns
            float act[kGColsPerWarp];
#pragma unroll
            for (int c = 0; c < kGColsPerWarp; ++c) act[c] = 0.f;
            for (int sg = 0; sg < nseg; ++sg, ++sgc) {
                const int ab = sgc & 1;
                mbar_wait(&bar_acc_full[ab], (sgc >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float da[16], db[16];
                    tmem_ld16(t_lane + ab * 256 + half * 16, da);
                    tmem_ld16(t_lane + ab * 256 + 128 + half * 16, db);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) act[half * 16 + c] += fmaf(db[c], 1.0f / kLoScale, da[c]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (DUO) mbar_arrive_cluster(mapa_u32(&bar_acc_empty[ab], 0));     // the leader's barrier
                    else mbar_arrive(&bar_acc_empty[ab]);
                }
            }
            unsigned badmask = 0;
            __half* r1row = a.R + (t < a.T ? t : 0) * a.Npr + col0;
            __half* r2row = r1row + a.plane;
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {                                      // 8 columns at a time
                float da[8], db[8];
                uint32_t h1[4], h2[4];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int cc = sub * 8 + c;
                    const float ism = __shfl_sync(0xffffffffu, ism_l, cc), bc = __shfl_sync(0xffffffffu, bias_l, cc);
                    float term = 0.f, r = 0.f;
                    if (col0 + cc < a.ncols && !(a.debug & 4)) {                     // warp-uniform
                        const float x = fmaf(act[cc], ism, bc);
                        if (NLIN == PYGLM_B200_NLIN_EXP && lv != 0.f && !(x <= kExpSafe)) badmask |= 1u << cc;
                        poisson_terms<NLIN>(x, (float)((sp[cc >> 2] >> ((cc & 3) * 8)) & 0xffu), a.dt, term, r);
                        term *= lv;
                        r *= lv;
                    }
                    da[c] = term;
                    db[c] = r;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float ra = db[2 * k] * kRScale, rb = db[2 * k + 1] * kRScale;
                    const __half2 hi = __floats2half2_rn(ra, rb);
                    const float2 back = __half22float2(hi);
                    const __half2 lo = __floats2half2_rn((ra - back.x) * kLoScale, (rb - back.y) * kLoScale);
                    h1[k] = *reinterpret_cast<const uint32_t*>(&hi);
                    h2[k] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                if (t < a.T) {
                    *reinterpret_cast<uint4*>(r1row + sub * 8) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                    *reinterpret_cast<uint4*>(r2row + sub * 8) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                }
                // column sums over this warp's 32 bins: lane l ends with column sub*8 + (l & 7)
                const float sl = warp_column_sums<8>(da, lane);
                const float sg = warp_column_sums<8>(db, lane);
                if (lane < 8) {                                                  // slot owned by this (CTA, quarter, lane); zeroed by the host
                    double* slot = my_part + (int64_t)(col0 + sub * 8 + lane) * 2;
                    slot[0] += (double)sl;
                    slot[1] += (double)sg;
                }
            }
            if (NLIN == PYGLM_B200_NLIN_EXP && a.flags) {
                badmask = __reduce_or_sync(0xffffffffu, badmask);
                if (((badmask >> lane) & 1u) && col0 + lane < a.ncols) a.flags[a.n_lo + col0 + lane] = 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) cluster_sync_all();         // neither CTA leaves while the other may still signal its barriers / TMEM
    if (warp == 0) {
        if (DUO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// gradient GEMM, split over time
// ---------------------------------------------------------------------------------------------
struct GemmBwdArgs {
    int64_t ntt;                       // 128-bin tiles in the recording
    int64_t tiles_per_split;
    int Npr; int64_t NBp;              // padded neurons / features (multiples of 128)
    double* Gp;                        // [splits][NBp][Npr]
};

template <bool WIDE>
__global__ void __launch_bounds__(kGThreads, 1)
tc_gemm_bwd_kernel(const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
                   const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ CUtensorMap mapR2, GemmBwdArgs a)
{
    // WIDE: operands as 64-element blocks in 128-byte rows (SWIZZLE_128B), two blocks per 128 features / neurons;
    // otherwise 32-element blocks in 64-byte rows (SWIZZLE_64B), four blocks.  Same bytes per stage either way.
    constexpr int kBlk = WIDE ? 64 : 32;                        // elements per block along M / N
    constexpr int kNBlk = 128 / kBlk;                           // blocks per operand plane
    constexpr int kBlkBytes = kBwdRows * kBlk * 2;              // one block: 64 bins x kBlk elements
    constexpr uint32_t kLayout = WIDE ? kSw128 : kSw64;
    constexpr uint32_t kSbo = 8 * kBlk * 2;                     // eight bins
    constexpr uint32_t kStepBytes = 16 * kBlk * 2;              // one k16 step: sixteen bins
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdStages * kBwdStageBytes);
    uint64_t* bar_full = bars;                       // [3]
    uint64_t* bar_empty = bars + kBwdStages;         // [3]
    uint64_t* bar_g_full = bars + 2 * kBwdStages;    // [2]
    uint64_t* bar_g_empty = bar_g_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_g_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fb = blockIdx.x, cb = blockIdx.y, split = blockIdx.z;
    const int64_t tile_lo = (int64_t)split * a.tiles_per_split;
    const int64_t tile_hi = min(a.ntt, tile_lo + a.tiles_per_split);
    const int nt = tile_hi > tile_lo ? (int)(tile_hi - tile_lo) : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kBwdStages; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_g_full[i], 1); mbar_init(&bar_g_empty[i], kGEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kGTmaWarp) {
        if (lane == 0) {
            for (int h = 0; h < 2 * nt; ++h) {                       // 64-bin half tiles
                const int s = h % kBwdStages;
                mbar_wait(&bar_empty[s], ((h / kBwdStages) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar_full[s], kBwdStageBytes);
                unsigned char* st = smem + s * kBwdStageBytes;
                const int row0 = (int)(tile_lo * kGT) + h * kBwdRows;
#pragma unroll
                for (int c = 0; c < kNBlk; ++c) {
                    tma_load_2d(st + (0 * kNBlk + c) * kBlkBytes, &mapX1, &bar_full[s], fb * kGF + c * kBlk, row0);
                    tma_load_2d(st + (1 * kNBlk + c) * kBlkBytes, &mapX2, &bar_full[s], fb * kGF + c * kBlk, row0);
                    tma_load_2d(st + (2 * kNBlk + c) * kBlkBytes, &mapR1, &bar_full[s], cb * kGN + c * kBlk, row0);
                    tma_load_2d(st + (3 * kNBlk + c) * kBlkBytes, &mapR2, &bar_full[s], cb * kGN + c * kBlk, row0);
                }
            }
        }
    } else if (warp == kGMmaWarp) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kGF, kGN, 1, 1);   // both operands MN-major
            constexpr uint32_t idesc2 = umma_idesc(kGF, 2 * kGN, 1, 1);   // r1 | r2 stacked (adjacent chunks, adjacent TMEM)
            for (int i = 0; i < nt; ++i) {
                const int gb = i & 1;
                const uint32_t t_a = tmem_base + gb * 256, t_b = t_a + 128;
                mbar_wait(&bar_g_empty[gb], ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int hh = 0; hh < 2; ++hh) {
                    const int h = 2 * i + hh, s = h % kBwdStages;
                    mbar_wait(&bar_full[s], (h / kBwdStages) & 1);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + s * kBwdStageBytes);
                    const uint64_t dx1 = umma_desc_layout(base, kBlkBytes, kSbo, kLayout);
                    const uint64_t dx2 = umma_desc_layout(base + 1 * kNBlk * kBlkBytes, kBlkBytes, kSbo, kLayout);
                    const uint64_t dr1 = umma_desc_layout(base + 2 * kNBlk * kBlkBytes, kBlkBytes, kSbo, kLayout);
#pragma unroll
                    for (int ks = 0; ks < kBwdRows / 16; ++ks) {
                        const uint32_t acc = (hh | ks) ? 1u : 0u;
                        const uint64_t off = (uint64_t)(ks * kStepBytes >> 4);
                        umma_f16(t_a, dx1 + off, dr1 + off, idesc2, acc);         // X1^T [r1 | r2]
                        umma_f16(t_b, dx2 + off, dr1 + off, idesc, 1u);           // X2^T r1
                    }
                    umma_commit(&bar_empty[s]);
                }
                umma_commit(&bar_g_full[gb]);
            }
        }
    } else {
        const int q = warp & 3;
        const int cg = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + cg * kGColsPerWarp;
        double gacc[kGColsPerWarp];
#pragma unroll
        for (int c = 0; c < kGColsPerWarp; ++c) gacc[c] = 0.0;
        for (int i = 0; i < nt; ++i) {
            const int gb = i & 1;
            mbar_wait(&bar_g_full[gb], (i >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
                float ga[8], gbv[8];
                tmem_ld8(t_lane + gb * 256 + sub * 8, ga);
                tmem_ld8(t_lane + gb * 256 + 128 + sub * 8, gbv);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) gacc[sub * 8 + c] += (double)fmaf(gbv[c], 1.0f / kLoScale, ga[c]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_g_empty[gb]);
        }
        double* dst = a.Gp + ((int64_t)split * a.NBp + (int64_t)fb * kGF + row) * a.Npr + cb * kGN + cg * kGColsPerWarp;
#pragma unroll
        for (int c = 0; c < kGColsPerWarp; ++c) dst[c] = gacc[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// final reductions (fixed order)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tc_gemm_final_ll_kernel(const double* __restrict__ part, int nctas, int Npr, int ncols,
                        double* __restrict__ out_ll, double* __restrict__ out_gb)
{
    // one block per column: 256 threads stride over the (CTA, quarter) slots, then a fixed-order tree
    __shared__ double sl[256], sg[256];
    const int nl = blockIdx.x;
    double l = 0.0, g = 0.0;
    for (int c = threadIdx.x; c < nctas * 4; c += 256) {
        l += part[((int64_t)c * Npr + nl) * 2];
        g += part[((int64_t)c * Npr + nl) * 2 + 1];
    }
    sl[threadIdx.x] = l; sg[threadIdx.x] = g;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) { sl[threadIdx.x] += sl[threadIdx.x + off]; sg[threadIdx.x] += sg[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out_ll[nl] = sl[0];
        if (out_gb) out_gb[nl] = sg[0];
    }
}

__global__ void __launch_bounds__(256)
tc_gemm_final_G_kernel(const double* __restrict__ Gp, int splits, int64_t NBp, int Npr, int N, int B, int F, int n_lo, int ncols,
                       const float* __restrict__ sx, const int8_t* __restrict__ A, const double* __restrict__ W,
                       double* __restrict__ out_gw)
{
    const int64_t NS = (int64_t)N * B, NB = NS + F;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (j, nl), nl fastest: coalesced partial reads
    if (idx >= NB * ncols) return;
    const int64_t j = idx / ncols;
    const int nl = (int)(idx - j * ncols);
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += Gp[((int64_t)z * NBp + j) * Npr + nl];
    const int n = n_lo + nl, pre = (int)(j / B);
    const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
    const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
    out_gw[(int64_t)nl * NB + j] = (a * ww) * s / ((double)sx[j] * (double)kRScale);
}

// ---------------------------------------------------------------------------------------------
int launch_tc_gemm_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (a.T <= 0 || a.ncols <= 0) return PYGLM_B200_OK;
    int rc = tc_ensure_planes(a, ws, stream);
    if (rc) return rc;
    if (!ws.gemm) ws.gemm = new GemmWorkspace();
    GemmWorkspace& g = *static_cast<GemmWorkspace*>(ws.gemm);
    const int NB = a.N * a.B + a.F;
    // Forward kernel variant.  3 (default): single CTAs, 64-feature K chunks in 128-byte rows (SWIZZLE_128B), three
    // 64 KB stages.  0: the same with 32-feature chunks in 64-byte rows (SWIZZLE_64B), seven 32 KB stages -- its MMAs
    // fetch their operands half as efficiently (C3 forward: 2.20 ms against 1.91 ms).  1 / 2: the cluster experiments.
    int mode = 3;
#if PYGLM_GEMM_EXPERIMENTS                                       // quarantined: build with -DPYGLM_GEMM_EXPERIMENTS=1 to select them
    if (const char* env = getenv("PYGLM_GEMM_MODE")) mode = atoi(env);
    if (mode < 0 || mode > 3) mode = 3;
#endif
    const int chunkf = mode == 3 ? 64 : 32;                      // features per K chunk
    const int nkc = (int)ceil_div(NB, chunkf);
    const int Kp = nkc * chunkf;
    const int Npr = (int)round_up(a.ncols, kGN);
    const int ncb = Npr / kGN;
    const int64_t ntt = ceil_div(a.T, kGT);
    const int64_t NBp = round_up(NB, kGF);
    const int nfb = (int)(NBp / kGF);
    const bool grad = a.out_gw != nullptr;

    if ((rc = ensure((void**)&g.Mp, &g.Mp_elems, (size_t)2 * Npr * Kp, sizeof(__half)))) return rc;
    if ((rc = ensure((void**)&g.colpar, &g.colpar_elems, (size_t)2 * Npr, sizeof(float)))) return rc;
    if ((rc = ensure((void**)&g.R, &g.R_elems, (size_t)2 * a.T * Npr, sizeof(__half)))) return rc;
    // Modes 1 (X multicast over a cluster of two) and 2 (cta_group::2 pairs) are parity-tested experiments on top of the
    // 32-feature layout; measured at C3 (forward only) when mode 0 took 2.47 ms: 2.6 / 2.75 ms.  With the MMAs removed the
    // TMA pipeline alone takes 1.7 ms (20 GB at the ~12 TB/s L2-to-SM cap) in mode 0 and 2.8 ms in mode 2: the pair's
    // cross-SM barrier traffic costs more than its 25 % fewer bytes save.
    if (mode == 1 && (ncb % 2 != 0)) mode = 0;
    if ((mode == 1 || mode == 2) && ws.num_sms < 2) mode = 0;
    const int64_t nwork = mode == 2 ? ceil_div(ntt, 2) * ncb * 2 : ntt * ncb;            // in CTAs
    int nctas = (int)std::min<int64_t>(nwork, ws.num_sms);
    if (mode == 1 || mode == 2) nctas &= ~1;
    if ((rc = ensure((void**)&g.part, &g.part_elems, (size_t)nctas * 4 * Npr * 2, sizeof(double)))) return rc;

    PYGLM_CUDA(cudaMemsetAsync(g.part, 0, (size_t)nctas * 4 * Npr * 2 * sizeof(double), stream));
    tc_gemm_prep_M_kernel<<<Npr, 256, 0, stream>>>(a.w, a.A, a.W, a.bias, ws.sx, a.N, a.B, a.F, a.n_lo, a.ncols, Npr, Kp, g.Mp, g.colpar);
    PYGLM_CUDA(cudaGetLastError());

    const CUtensorMap* xmaps = static_cast<const CUtensorMap*>(ws.tmaps);
    CUtensorMap mM1, mM2, mR1, mR2;
    if ((rc = tc_make_map_2d(&mM1, g.Mp, Kp, Npr, Kp, chunkf, kGN, mode == 3))) return rc;
    if ((rc = tc_make_map_2d(&mM2, g.Mp + (size_t)Npr * Kp, Kp, Npr, Kp, chunkf, kGN, mode == 3))) return rc;

    GemmFwdArgs f{};
    f.flags = a.flags;
    f.Sp = ws.Sp; f.Nps = ws.Np; f.T = a.T; f.n_lo = a.n_lo; f.ncols = a.ncols; f.Npr = Npr; f.nkc = nkc;
    // up to ~1500 features one segment is accurate enough (measured: 7e-7 on ll, 2e-6 on gradients at 1280 features,
    // exp model) and costs nothing; beyond that the K loop is cut into 256-feature segments (~9 % slower, 70x more accurate
    // at 10240 features)
    f.seg = nkc * chunkf <= kFwdSingleSegmentChunks * 32 ? nkc : kFwdSegChunks * 32 / chunkf;
    if (const char* env = getenv("PYGLM_GEMM_DEBUG")) f.debug = atoi(env);
    if (const char* env = getenv("PYGLM_GEMM_SEG")) f.seg = std::max(1, atoi(env));      // precision experiments
    f.ntt = ntt; f.ncb = ncb; f.colpar = g.colpar; f.R = g.R; f.plane = (int64_t)a.T * Npr; f.part = g.part; f.dt = (float)a.dt;
    const int smem_f = kFwdStages * kFwdStageBytes + 256 + 1024;
    if (!g.mapX64_ready) {
        if ((rc = tc_make_map_2d(&g.mapX64[0], ws.X1, ws.ldp, a.T, ws.ldp, 32, kBwdRows))) return rc;
        if ((rc = tc_make_map_2d(&g.mapX64[1], ws.X2, ws.ldp, a.T, ws.ldp, 32, kBwdRows))) return rc;
        g.mapX64_ready = true;
    }
    if (mode == 3) {
        CUtensorMap mX1w, mX2w;                                   // 64 features x 128 bins boxes, 128-byte swizzle
        if ((rc = tc_make_map_2d(&mX1w, ws.X1, ws.ldp, a.T, ws.ldp, 64, kGT, true))) return rc;
        if ((rc = tc_make_map_2d(&mX2w, ws.X2, ws.ldp, a.T, ws.ldp, 64, kGT, true))) return rc;
        auto kf = a.nlin == PYGLM_B200_NLIN_EXP ? tc_gemm_fwd_kernel<PYGLM_B200_NLIN_EXP, 3>
                                                : tc_gemm_fwd_kernel<PYGLM_B200_NLIN_SOFTPLUS, 3>;
        PYGLM_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
        kf<<<nctas, kGThreads, smem_f, stream>>>(mX1w, mX2w, mM1, mM2, f);
    }
#if PYGLM_GEMM_EXPERIMENTS
    else if (mode) {
        CUtensorMap mM1h, mM2h;                                   // 64-column boxes of the weight planes (mode 2)
        if ((rc = tc_make_map_2d(&mM1h, g.Mp, Kp, Npr, Kp, 32, kGN / 2))) return rc;
        if ((rc = tc_make_map_2d(&mM2h, g.Mp + (size_t)Npr * Kp, Kp, Npr, Kp, 32, kGN / 2))) return rc;
        void (*kf)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmFwdArgs);
        if (mode == 2) kf = a.nlin == PYGLM_B200_NLIN_EXP ? tc_gemm_fwd_kernel<PYGLM_B200_NLIN_EXP, 2> : tc_gemm_fwd_kernel<PYGLM_B200_NLIN_SOFTPLUS, 2>;
        else kf = a.nlin == PYGLM_B200_NLIN_EXP ? tc_gemm_fwd_kernel<PYGLM_B200_NLIN_EXP, 1> : tc_gemm_fwd_kernel<PYGLM_B200_NLIN_SOFTPLUS, 1>;
        PYGLM_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)nctas); cfg.blockDim = dim3(kGThreads); cfg.dynamicSmemBytes = smem_f; cfg.stream = stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (mode == 2) PYGLM_CUDA(cudaLaunchKernelEx(&cfg, kf, xmaps[0], xmaps[1], mM1h, mM2h, f));
        else PYGLM_CUDA(cudaLaunchKernelEx(&cfg, kf, g.mapX64[0], g.mapX64[1], mM1, mM2, f));
    } else {
        auto kf = a.nlin == PYGLM_B200_NLIN_EXP ? tc_gemm_fwd_kernel<PYGLM_B200_NLIN_EXP, 0>
                                                : tc_gemm_fwd_kernel<PYGLM_B200_NLIN_SOFTPLUS, 0>;
        PYGLM_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
        kf<<<nctas, kGThreads, smem_f, stream>>>(xmaps[0], xmaps[1], mM1, mM2, f);
    }
#endif
    PYGLM_CUDA(cudaGetLastError());
    tc_gemm_final_ll_kernel<<<(unsigned)a.ncols, 256, 0, stream>>>(g.part, nctas, Npr, a.ncols, a.out_ll, a.out_gb);
    PYGLM_CUDA(cudaGetLastError());
    if (!grad) return PYGLM_B200_OK;

    bool bwd_wide = true;                                        // 128-byte-swizzled operand blocks (64 elements)
    if (const char* env = getenv("PYGLM_GEMM_BWD_WIDE")) bwd_wide = atoi(env) != 0;
    const int blk = bwd_wide ? 64 : 32;
    if ((rc = tc_make_map_2d(&mR1, g.R, Npr, a.T, Npr, blk, kBwdRows, bwd_wide))) return rc;
    if ((rc = tc_make_map_2d(&mR2, g.R + (size_t)a.T * Npr, Npr, a.T, Npr, blk, kBwdRows, bwd_wide))) return rc;
    CUtensorMap mXb1, mXb2;
    if ((rc = tc_make_map_2d(&mXb1, ws.X1, ws.ldp, a.T, ws.ldp, blk, kBwdRows, bwd_wide))) return rc;
    if ((rc = tc_make_map_2d(&mXb2, ws.X2, ws.ldp, a.T, ws.ldp, blk, kBwdRows, bwd_wide))) return rc;
    int64_t splits = std::max<int64_t>(1, (int64_t)ws.num_sms / ((int64_t)nfb * ncb));
    splits = std::min<int64_t>(splits, ntt);
    const int64_t tps = ceil_div(ntt, splits);
    splits = ceil_div(ntt, tps);
    if ((rc = ensure((void**)&g.Gp, &g.Gp_elems, (size_t)splits * NBp * Npr, sizeof(double)))) return rc;
    GemmBwdArgs b{};
    b.ntt = ntt; b.tiles_per_split = tps; b.Npr = Npr; b.NBp = NBp; b.Gp = g.Gp;
    const int smem_b = kBwdStages * kBwdStageBytes + 256 + 1024;
    auto kb = bwd_wide ? tc_gemm_bwd_kernel<true> : tc_gemm_bwd_kernel<false>;
    PYGLM_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
    dim3 gridb((unsigned)nfb, (unsigned)ncb, (unsigned)splits);
    kb<<<gridb, kGThreads, smem_b, stream>>>(mXb1, mXb2, mR1, mR2, b);
    PYGLM_CUDA(cudaGetLastError());
    tc_gemm_final_G_kernel<<<(unsigned)ceil_div((int64_t)NB * a.ncols, 256), 256, 0, stream>>>(
        g.Gp, (int)splits, NBp, Npr, a.N, a.B, a.F, a.n_lo, a.ncols, ws.sx, a.A, a.W, a.out_gw);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

}  // namespace pyglm
