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

#include "allreduce.cuh"
#include "llgrad_tc.cuh"
#include "tc_common.cuh"

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

#ifndef PYGLM_TC_COLGROUPS
#define PYGLM_TC_COLGROUPS 4
#endif
constexpr int kColGroups = PYGLM_TC_COLGROUPS;  // column groups: the epilogue runs 4 * kColGroups warps
constexpr int kColsPerWarp = kNcol / kColGroups;
constexpr int kEpiWarps = 4 * kColGroups;
#ifndef PYGLM_TC_ROLE_BASE
#define PYGLM_TC_ROLE_BASE 1
#endif
// Warp roles.  Default: epilogue warps 0..15 first, then an idle warp and the TMA producer / forward-MMA / gradient-MMA
// issuers on schedulers 1..3, so no spinning role warp shares a scheduler slot pattern with lane quarter 0 (measured:
// with the role warps first, the quarter-0 epilogue warps trailed the others by ~1.3k cycles per tile).
#if PYGLM_TC_ROLE_BASE == 0
constexpr int kFirstEpiWarp = 3, kProducerWarp = 0, kFwdWarp = 1, kBwdWarp = 2;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);
#elif PYGLM_TC_ROLE_BASE == 2
// no idle warp: 19 warps leave 104 registers per thread instead of 96
constexpr int kFirstEpiWarp = 0, kProducerWarp = kEpiWarps, kFwdWarp = kEpiWarps + 1, kBwdWarp = kEpiWarps + 2;
constexpr int kThreads = 32 * (kEpiWarps + 3);
#else
constexpr int kFirstEpiWarp = 0, kProducerWarp = kEpiWarps + 1, kFwdWarp = kEpiWarps + 2, kBwdWarp = kEpiWarps + 3;
constexpr int kThreads = 32 * (kEpiWarps + 4);
#endif
constexpr int kFlushTiles = 8;              // TMEM gradient accumulators are folded into FP64 every 8 tiles

void TcWorkspace::release()
{
    cudaFree(Sp); Sp = nullptr;
    cudaFree(acc); acc = nullptr; acc_elems = 0; streamed = false; chunk_rows = 0; map_rows = 0;
    cudaFree(colflag); colflag = nullptr;
    cudaFree(X1); cudaFree(X2); cudaFree(sx); cudaFree(colmax); cudaFree(Mp); cudaFree(colpar); cudaFree(part);
    X1 = X2 = nullptr; sx = nullptr; colmax = nullptr; Mp = nullptr; colpar = nullptr; part = nullptr;
    part_elems = 0; planes_ready = false;
    if (tmaps) free(tmaps);
    tmaps = nullptr;
    tc_gemm_release(gemm);
    gemm = nullptr;
}

bool tc_supported(int64_t T, int N, int B, int x_dtype)
{
    return x_dtype == PYGLM_B200_X_F32 && T > 0;
}

// the fused kernel's final reduction can carry the sum over ranks when one launch covers the population (V2 kernel)
bool tc_can_fuse_allreduce(int N, int64_t nfeat) { return N <= kNcol && nfeat > 3 * kChunkF && nfeat <= kMaxChunks * kChunkF && nfeat + 1 <= kArMaxBlocks; }

// N*B <= 160 features: the fused single-pass kernel; larger populations: the two GEMM kernels
bool tc_uses_fused_kernel(int64_t nfeat) { return nfeat <= kMaxChunks * kChunkF; }

// ---------------------------------------------------------------------------------------------
// One-time: per-feature scales and the FP16 split planes of X
// ---------------------------------------------------------------------------------------------
// Spike-history features: analytic bound |X[t][pre*B+b]| <= (largest count in column pre) * sum_k |ibasis[k][b]|, so the
// scales need no pass over X (and are the same whether the planes come straight from K1 or from a resident X).
__global__ void __launch_bounds__(256)
tc_spike_colmax_kernel(const uint8_t* __restrict__ S, int64_t nbytes, int N, unsigned* __restrict__ cmax)
{
    extern __shared__ unsigned s_cmax[];                                  // [N]
    for (int n = threadIdx.x; n < N; n += blockDim.x) s_cmax[n] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t mis = (int64_t)((16 - (reinterpret_cast<uintptr_t>(S) & 15)) & 15);
    const int64_t head = nbytes < mis ? nbytes : mis;
    const int64_t nvec = (nbytes - head) >> 4;
    auto note = [&](int64_t f, unsigned v) {
        const int n = (int)(f % N);
        if (v > s_cmax[n]) atomicMax(&s_cmax[n], v);
    };
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid < head && S[gtid]) note(gtid, S[gtid]);
    const uint4* vp = reinterpret_cast<const uint4*>(S + head);
    for (int64_t v = gtid; v < nvec; v += stride) {
        const uint4 q = __ldg(vp + v);
        if ((q.x | q.y | q.z | q.w) == 0u) continue;
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                const unsigned b = (w[i] >> (8 * j)) & 0xffu;
                if (b) note(head + 16 * v + 4 * i + j, b);
            }
    }
    for (int64_t f = head + 16 * nvec + gtid; f < nbytes; f += stride) if (S[f]) note(f, S[f]);
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) if (s_cmax[n]) atomicMax(&cmax[n], s_cmax[n]);
}

__global__ void tc_spike_scales_kernel(const unsigned* __restrict__ cmax, const double* __restrict__ ibasis, int R, int B, int N,
                                       float* __restrict__ sx)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N * B) return;
    const int pre = j / B, b = j - pre * B;
    double s = 0.0;
    for (int k = 0; k < R; ++k) s += fabs(ibasis[(size_t)k * B + b]);
    sx[j] = pow2_scale((float)(s * (double)cmax[pre]) * 1.0001f);
}

// Stimulus features (real valued): the scale comes from the data.
__global__ void tc_colmax_kernel(const float* __restrict__ X, int64_t T, int j0, int NB, int64_t ldx, unsigned* colmax)
{
    const int j = j0 + blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= NB) return;
    float m = 0.f;
    for (int64_t t = blockIdx.x; t < T; t += gridDim.x) m = fmaxf(m, fabsf(X[t * ldx + j]));
    atomicMax(&colmax[j], __float_as_uint(m));       // non-negative floats order like unsigned ints
}

__global__ void tc_scales_kernel(const unsigned* colmax, int j0, int NB, float* sx)
{
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
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
                 int N, int B, int F, int n_lo, int ncols, int Kp, __half* __restrict__ Mp, float* __restrict__ colpar)
{
    __shared__ float smax[256];
    const int nl = blockIdx.x;                 // 0..31
    const int NS = N * B, NB = NS + F;         // spike-history features, all features
    const bool live = nl < ncols;
    const int n = n_lo + nl;
    float mx = 0.f;
    if (live) {
        for (int j = threadIdx.x; j < NB; j += 256) {
            const int pre = j / B;
            const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
            const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
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
        colpar[kNcol + nl] = live ? (float)bias[n] : 0.f;
    }
}

struct TcKernelArgs {
    const uint8_t* S; int64_t T; int N; int halo;
    const uint8_t* Sp; int Np;     // padded copy of the spikes [T][Np], Np % 32 == 0 (vector loads)
    float dt; int nlin; int n_lo, ncols;
    int nch;                       // 32-feature chunks (1..5)
    int nmt;                       // 128-feature gradient tiles (1..2)
    int64_t ntiles;
    const __half* Mp; int Kp;      // [2][32][Kp]
    const float* colpar;           // [2][32]
    double* part;                  // per CTA: [nmt][128][32] G, then [32] ll, [32] g_bias
    int nfeat;                     // N*B real features
    int flush;                     // fold the TMEM gradient accumulators into FP64 every `flush` tiles
    int debug;                     // timing experiments only (PYGLM_TC_DEBUG): skip phases
    unsigned producer_sleep_ns;
    long long* trace;              // optional [3][32][4] clock64 stamps of CTA 0 (PYGLM_TC_TRACE)
    unsigned* flags;               // [N] range flags, or nullptr
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
    L.off_bar = L.off_R + 2 * 2 * kRBytes;            // residual planes are double-buffered
    // slack so the second gradient tile's descriptor (4 chunks from chunk 4) stays inside the allocation
    int end = L.off_bar + 512;
    const int reach = (kStages - 1) * L.stage_bytes + nch * kChunkBytes + 8 * kChunkBytes;
    L.total = end > reach ? end : reach;
    return L;
}

// TMEM columns: forward accumulators at 0: [X1 M1 | X1 M2 | X2 M1]; gradient buffer gb=0,1, tile mt=0,1 at
// 96 + gb*192 + mt*96: [X1^T r1 | X1^T r2 | X2^T r1].  The second and third block of each carry 2^-11.
constexpr int kFwdCols = 96;
constexpr int kGradBase = kFwdCols;          // forward accumulators are single-buffered (drained at once)
constexpr int kGradBuf = 2 * kFwdCols;       // one gradient buffer = two 128-feature tiles; double-buffered
constexpr int kTmemAlloc = 512;

// WIDE (4 or 5 chunks, i.e. 97..160 features): features 0..127 travel as two 64-feature chunks in 128-byte rows
// (SWIZZLE_128B) -- the tensor core fetches those with full shared-memory wavefronts, 64-byte-swizzled operands with
// half-empty ones -- and only the tail chunk (features 128..159) keeps the 32-feature / 64-byte layout.  Stage size,
// TMEM layout and epilogue are the same in both variants.
template <int NLIN, bool WIDE>
__global__ void __launch_bounds__(kThreads, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap tmap1, const __grid_constant__ CUtensorMap tmap2,
                const __grid_constant__ CUtensorMap tmapW1, const __grid_constant__ CUtensorMap tmapW2, TcKernelArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const TcSmem L = tc_smem_layout(a.nch);
    unsigned char* sM = smem + L.off_M;               // per chunk: [M1 rows 0..31][M2 rows 0..31], 64 B rows
    unsigned char* sR = smem + L.off_R;               // buffer b: [r1 plane][r2 plane]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* bar_full = bars;            // [2 stages][5 chunks] TMA landed, per 32-feature chunk (both planes)
    uint64_t* bar_empty = bars + 10;      // [2 stages][2 groups] chunks 0-3 / chunk 4 consumed by the gradient MMA
    uint64_t* bar_fwd_full = bars + 14;   // activation accumulators ready
    uint64_t* bar_fwd_empty = bars + 15;  // ... drained by the epilogue
    uint64_t* bar_r_ready = bars + 16;    // [2] residual planes written
    uint64_t* bar_r_free = bars + 18;     // [2] ... consumed by the gradient MMA
    uint64_t* bar_g_full = bars + 20;     // [2] gradient accumulators of a tile complete
    uint64_t* bar_g_empty = bars + 22;    // [2] ... folded into FP64
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    float* sPar = reinterpret_cast<float*>(bars + 26);   // [32] x (1/sm, bias)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = a.nch;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            for (int c = 0; c < kMaxChunks; ++c) mbar_init(&bar_full[i * kMaxChunks + c], 1);
            mbar_init(&bar_empty[2 * i], 1); mbar_init(&bar_empty[2 * i + 1], 1);
            mbar_init(&bar_g_full[i], 1); mbar_init(&bar_g_empty[i], kEpiWarps);
            mbar_init(&bar_r_ready[i], kEpiWarps); mbar_init(&bar_r_free[i], 1);
        }
        mbar_init(bar_fwd_full, 1); mbar_init(bar_fwd_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemAlloc) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // weight planes -> shared memory: K-major rows of 64 B, 64B swizzle applied by hand; within a
    // chunk the 32 rows of M1 are followed by the 32 rows of M2 so one N=64 MMA sees [M1 | M2]
    {
        const int n16 = 2 * nch * kMChunkBytes / 16;         // 16-byte units over both planes
        for (int u = threadIdx.x; u < n16; u += kThreads) {
            const int c = u / (2 * kMChunkBytes / 16);       // chunk
            const int v = u - c * (2 * kMChunkBytes / 16);
            const int plane = v / (kMChunkBytes / 16);
            const int w = v - plane * (kMChunkBytes / 16);
            const int row = w >> 2, q = w & 3;               // row n (64 B), 16-byte unit within the row
            const uint4 val = *reinterpret_cast<const uint4*>(
                a.Mp + ((int64_t)(plane * kNcol + row) * a.Kp + c * kChunkF + q * 8));
            if (WIDE && c < 4) {
                // wide chunk c/2: 64 rows [M1 | M2] of 128 B, this narrow chunk is its half (c & 1); 128-byte swizzle
                const int r = plane * kNcol + row, q8 = (c & 1) * 4 + q;
                *reinterpret_cast<uint4*>(sM + (c >> 1) * 4 * kMChunkBytes + r * 128 + ((q8 ^ (r & 7)) << 4)) = val;
            } else {
                *reinterpret_cast<uint4*>(sM + c * 2 * kMChunkBytes + plane * kMChunkBytes + row * 64 + ((q ^ ((row >> 1) & 3)) << 4)) = val;
            }
        }
        fence_proxy_async();
        if (threadIdx.x < kNcol) {
            sPar[2 * threadIdx.x] = a.colpar[threadIdx.x];
            sPar[2 * threadIdx.x + 1] = a.colpar[kNcol + threadIdx.x];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, step = gridDim.x;
    const int ntl = first < a.ntiles ? (int)((a.ntiles - first + step - 1) / step) : 0;   // tiles of this CTA
    const int F = a.flush;

    if (warp == kProducerWarp) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            for (int it = 0; it < ntl; ++it) {
                const int s = it & 1;
                // Chunks land on their own barriers so the forward MMA starts on chunk 0 while chunk 4 is still in
                // flight, and chunks 0-3 are refilled as soon as the first gradient tile has consumed them.
                unsigned char* st = smem + s * L.stage_bytes;
                const int row0 = (int)((first + (int64_t)it * step) * kTileT);
                const uint32_t par = ((it >> 1) & 1) ^ 1;
                if constexpr (WIDE) {
                    // load units: two wide chunks (64 features x 128 bins per plane), then the tail chunk if there is one
                    const int nu = 2 + (nch - 4);
                    for (int u = 0; u < nu; ++u) {
                        if (u == 0) mbar_wait_relaxed(&bar_empty[2 * s], par, a.producer_sleep_ns);
                        if (u == 2) mbar_wait_relaxed(&bar_empty[2 * s + 1], par, a.producer_sleep_ns);
                        uint64_t* fb = &bar_full[s * kMaxChunks + u];
                        if (u == 0 && a.trace && blockIdx.x == 0 && it < 32) a.trace[(0 * 32 + it) * 4 + 0] = clock64();
                        if (u < 2) {
                            mbar_arrive_expect_tx(fb, 4 * kChunkBytes);
                            tma_load_2d(st + u * 2 * kChunkBytes, &tmapW1, fb, u * 64, row0);
                            tma_load_2d(st + (nch + u * 2) * kChunkBytes, &tmapW2, fb, u * 64, row0);
                        } else {
                            mbar_arrive_expect_tx(fb, 2 * kChunkBytes);
                            tma_load_2d(st + 4 * kChunkBytes, &tmap1, fb, 4 * kChunkF, row0);
                            tma_load_2d(st + (nch + 4) * kChunkBytes, &tmap2, fb, 4 * kChunkF, row0);
                        }
                    }
                    continue;
                }
                for (int c = 0; c < nch; ++c) {
                    if (c == 0) mbar_wait_relaxed(&bar_empty[2 * s], par, a.producer_sleep_ns);
                    if (c == 4) mbar_wait_relaxed(&bar_empty[2 * s + 1], par, a.producer_sleep_ns);
                    uint64_t* fb = &bar_full[s * kMaxChunks + c];
                    if ((a.debug & 16) && it >= 2) { mbar_arrive(fb); continue; }
                    if (c == 0 && a.trace && blockIdx.x == 0 && it < 32) a.trace[(0 * 32 + it) * 4 + 0] = clock64();
                    mbar_arrive_expect_tx(fb, 2 * kChunkBytes);
                    tma_load_2d(st + c * kChunkBytes, &tmap1, fb, c * kChunkF, row0);
                    tma_load_2d(st + (nch + c) * kChunkBytes, &tmap2, fb, c * kChunkF, row0);
                }
            }
        }
    } else if (warp == kFwdWarp) {
        // ================================ forward MMA issuer ==========================
        if (lane == 0) {
            constexpr uint32_t idesc_f64 = umma_idesc(128, 64, 0, 0);     // A: X K-major,  B: [M1|M2] K-major
            constexpr uint32_t idesc_f32 = umma_idesc(128, 32, 0, 0);
            const uint32_t sMb = smem_u32(sM);
            const uint32_t t_f = tmem_base;
            for (int it = 0; it < ntl; ++it) {
                const int s = it & 1;
                const uint32_t sX1 = smem_u32(smem + s * L.stage_bytes), sX2 = sX1 + nch * kChunkBytes;
                mbar_wait(bar_fwd_empty, (it & 1) ^ 1);
                if constexpr (WIDE) {
                    for (int u = 0; u < 2; ++u) {                                  // wide chunks: four k16 steps each
                        mbar_wait(&bar_full[s * kMaxChunks + u], (it >> 1) & 1);
                        if (u == 0 && a.trace && blockIdx.x == 0 && it < 32) a.trace[(1 * 32 + it) * 4 + 0] = clock64();
                        tc_fence_after();
                        const uint64_t wx1 = umma_desc_sw128(sX1 + u * 2 * kChunkBytes);
                        const uint64_t wx2 = umma_desc_sw128(sX2 + u * 2 * kChunkBytes);
                        const uint64_t wm = umma_desc_sw128(sMb + u * 4 * kMChunkBytes);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t acc = (u | ks) ? 1u : 0u;
                            umma_f16(t_f, wx1 + 2 * ks, wm + 2 * ks, idesc_f64, acc);          // X1 [M1 | M2]
                            umma_f16(t_f + 64, wx2 + 2 * ks, wm + 2 * ks, idesc_f32, acc);     // X2 M1
                        }
                    }
                    if (nch > 4) {                                                 // tail chunk, 64-byte layout
                        mbar_wait(&bar_full[s * kMaxChunks + 2], (it >> 1) & 1);
                        tc_fence_after();
                        const uint64_t dx1 = umma_desc(sX1 + 4 * kChunkBytes, 16, 512), dx2 = umma_desc(sX2 + 4 * kChunkBytes, 16, 512);
                        const uint64_t dm = umma_desc(sMb + 8 * kMChunkBytes, 16, 512);
                        umma_f16(t_f, dx1, dm, idesc_f64, 1u);
                        umma_f16(t_f + 64, dx2, dm, idesc_f32, 1u);
                        umma_f16(t_f, dx1 + 2, dm + 2, idesc_f64, 1u);
                        umma_f16(t_f + 64, dx2 + 2, dm + 2, idesc_f32, 1u);
                    }
                } else {
                uint64_t dx1 = umma_desc(sX1, 16, 512), dx2 = umma_desc(sX2, 16, 512), dm = umma_desc(sMb, 16, 512);
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(&bar_full[s * kMaxChunks + c], (it >> 1) & 1);
                    if (c == 0 && a.trace && blockIdx.x == 0 && it < 32) a.trace[(1 * 32 + it) * 4 + 0] = clock64();
                    tc_fence_after();
                    if (a.debug & 1) continue;
                    umma_f16(t_f, dx1, dm, idesc_f64, c ? 1u : 0u);            // X1 [M1 | M2], features 0..15 of the chunk
                    umma_f16(t_f + 64, dx2, dm, idesc_f32, c ? 1u : 0u);       // X2 M1
                    umma_f16(t_f, dx1 + 2, dm + 2, idesc_f64, 1u);             // features 16..31 (+32 bytes)
                    umma_f16(t_f + 64, dx2 + 2, dm + 2, idesc_f32, 1u);
                    dx1 += kChunkBytes >> 4; dx2 += kChunkBytes >> 4; dm += (2 * kMChunkBytes) >> 4;
                }
                }
                umma_commit(bar_fwd_full);
                if (a.trace && blockIdx.x == 0 && it < 32) a.trace[(1 * 32 + it) * 4 + 1] = clock64();
            }
        }
    } else if (warp == kBwdWarp) {
        // ================================ gradient MMA issuer =========================
        // Its own warp, so a forward MMA waiting for data never delays a gradient MMA (or the reverse).
        // The TMEM accumulators do not keep full FP32 precision over long sums (measured: the gradient
        // error grows linearly with the number of accumulated tiles), so every tile starts from zero
        // in one of two accumulator buffers and the epilogue folds it into FP64.
        if (lane == 0) {
            constexpr uint32_t idesc_b64 = umma_idesc(128, 64, 1, 1);     // A: X MN-major, B: [r1|r2] MN-major
            constexpr uint32_t idesc_b32 = umma_idesc(128, 32, 1, 1);
            for (int j = 0; j < ntl; ++j) {
                const int s = j & 1, b = j & 1, gb = j & 1;
                const uint32_t sX1 = smem_u32(smem + s * L.stage_bytes), sX2 = sX1 + nch * kChunkBytes;
                const uint32_t sR1 = smem_u32(sR + b * 2 * kRBytes);
                for (int c = 0; c < (WIDE ? 2 + (nch - 4) : nch); ++c) mbar_wait(&bar_full[s * kMaxChunks + c], (j >> 1) & 1);   // landed long ago
                mbar_wait(&bar_r_ready[b], (j >> 1) & 1);
                mbar_wait(&bar_g_empty[gb], ((j >> 1) & 1) ^ 1);
                if (a.trace && blockIdx.x == 0 && j < 32) a.trace[(1 * 32 + j) * 4 + 2] = clock64();
                tc_fence_after();
                // descriptors differ between k steps only in the start address: add (bytes >> 4)
                const uint64_t dr0 = umma_desc(sR1, kRBytes, 512);                // LBO steps r1 -> r2
                for (int mt = 0; mt < ((a.debug & 2) ? 0 : a.nmt); ++mt) {
                    const uint32_t t_g = tmem_base + kGradBase + gb * kGradBuf + mt * kFwdCols;
                    // Tile 0 spans chunks 0..3 (LBO = one chunk).  Tile 1 has a single chunk (features 128..159): with
                    // LBO = 0 all four 32-row groups of its M = 128 alias that chunk, so every TMEM lane quarter holds
                    // a copy of the same 32 gradient rows and the epilogue warps can take turns folding them.
                    const uint32_t lbo = mt ? 0u : (uint32_t)kChunkBytes;
                    const bool wide_tile = WIDE && mt == 0;       // two 64-feature blocks in 128-byte rows: LBO = one wide chunk
                    const uint64_t dx1_0 = wide_tile ? umma_desc_layout(sX1, 2 * kChunkBytes, 1024, kSw128)
                                                     : umma_desc(sX1 + mt * 4 * kChunkBytes, lbo, 512);
                    const uint64_t dx2_0 = wide_tile ? umma_desc_layout(sX2, 2 * kChunkBytes, 1024, kSw128)
                                                     : umma_desc(sX2 + mt * 4 * kChunkBytes, lbo, 512);
                    const uint32_t xstep = wide_tile ? 2048u : 1024u;             // sixteen bins of the X operand
#pragma unroll
                    for (int ks = 0; ks < kTileT / 16; ++ks) {                    // 16 bins per step
                        const uint32_t acc = ks ? 1u : 0u;
                        const uint64_t off = (uint64_t)(ks * 1024 >> 4), offx = (uint64_t)(ks * xstep >> 4);
                        umma_f16(t_g, dx1_0 + offx, dr0 + off, idesc_b64, acc);      // X1^T [r1 | r2]
                        umma_f16(t_g + 64, dx2_0 + offx, dr0 + off, idesc_b32, acc); // X2^T r1
                    }
                    if (mt == 0) umma_commit(&bar_empty[2 * s]);                 // chunks 0-3 can be refilled
                }
                if (a.debug & 2) umma_commit(&bar_empty[2 * s]);
                umma_commit(&bar_empty[2 * s + 1]);
                umma_commit(&bar_r_free[b]);
                umma_commit(&bar_g_full[gb]);
                if (a.trace && blockIdx.x == 0 && j < 32) a.trace[(1 * 32 + j) * 4 + 3] = clock64();
            }
        }
    } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kEpiWarps) {
        // ================================ epilogue warps ==============================
        // kEpiWarps warps; warp w owns TMEM lane quarter q = w & 3 (hardware restriction) and the
        // column group cg of kColsPerWarp postsynaptic columns.
        const int q = warp & 3;
        const int cg = (warp - kFirstEpiWarp) >> 2;
        const int c0 = cg * kColsPerWarp;                // first column of this warp
        const int row = q * 32 + lane;                   // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + c0;
        const float2* cpar = reinterpret_cast<const float2*>(sPar);     // (1/sm, bias) per column
        double* gp = a.part + (int64_t)blockIdx.x * ((int64_t)a.nmt * 128 * kNcol + 2 * kNcol);
        double gacc[kColsPerWarp], gacc1[kColsPerWarp];  // gradient tiles 0 / 1 (features 0..127 / 128..): FP64 in registers
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) { gacc[c] = 0.0; gacc1[c] = 0.0; }
        // fold tile j's TMEM gradient block into FP64.  Gradient tile 1 (features >= 128: one 32-feature chunk,
        // replicated into all four lane quarters by the issuer) is folded by quarter j % 4 only, so the extra work
        // rotates over the warps instead of doubling the load of quarter 0; tc_final_kernel adds the four copies.
        auto fold_gradient = [&](int j) {
            float g0[kColsPerWarp], g1[kColsPerWarp], g2[kColsPerWarp];
            const int gb = j & 1;
            const uint32_t t_g = t_lane + kGradBase + gb * kGradBuf;
            const bool trf = a.trace && blockIdx.x == 0 && warp == kFirstEpiWarp + (a.debug >> 8) && lane == 0 && j < 31;
            if (trf) a.trace[1536 + j * 4 + 0] = clock64();
            mbar_wait(&bar_g_full[gb], (j >> 1) & 1);
            if (trf) a.trace[1536 + j * 4 + 1] = clock64();
            tc_fence_after();
            if (!(a.debug & 8)) {
                const bool tile1 = a.nmt > 1 && (j & 3) == q;                  // warp-uniform
                tmem_ld<kColsPerWarp>(t_g + 0, g0);
                tmem_ld<kColsPerWarp>(t_g + 32, g1);
                tmem_ld<kColsPerWarp>(t_g + 64, g2);
                tmem_ld_wait();
                if (trf) a.trace[1536 + j * 4 + 2] = clock64();
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) gacc[c] += (double)fmaf(g1[c] + g2[c], 1.0f / kLoScale, g0[c]);
                if (tile1) {
                    tmem_ld<kColsPerWarp>(t_g + kFwdCols + 0, g0);
                    tmem_ld<kColsPerWarp>(t_g + kFwdCols + 32, g1);
                    tmem_ld<kColsPerWarp>(t_g + kFwdCols + 64, g2);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < kColsPerWarp; ++c) gacc1[c] += (double)fmaf(g1[c] + g2[c], 1.0f / kLoScale, g0[c]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_g_empty[gb]);
            if (trf) a.trace[1536 + j * 4 + 3] = clock64();
        };
        double ll_acc = 0.0, gb_acc = 0.0;               // lane l accumulates column c0 + (l % kColsPerWarp)
        const int sw = (row >> 1) & 3;

        // per-thread FP32 partial column sums, folded (butterfly + FP64) at the flush points
        float pll[kColsPerWarp], pgb[kColsPerWarp];
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) { pll[c] = 0.f; pgb[c] = 0.f; }

        // spikes of one bin for this warp's columns (kColsPerWarp bytes packed in registers); the next
        // tile's are fetched a tile ahead.  Aligned column ranges use one vector load from the padded copy.
        constexpr int kSpWords = kColsPerWarp / 4;
        uint32_t sb[kSpWords], sb_next[kSpWords];
        const bool vec_ok = ((a.n_lo + c0) % kColsPerWarp) == 0;
        auto load_spikes = [&](int it, uint32_t (&dst)[kSpWords]) {
            const int64_t t = (first + (int64_t)it * step) * kTileT + row;
            const bool ok = it < ntl && t < a.T;
#pragma unroll
            for (int i = 0; i < kSpWords; ++i) dst[i] = 0;
            if (!ok) return;
            if (vec_ok) {
                const uint8_t* src = a.Sp + t * a.Np + a.n_lo + c0;
                if constexpr (kSpWords == 2) { const uint2 v = *reinterpret_cast<const uint2*>(src); dst[0] = v.x; dst[1] = v.y; }
                else if constexpr (kSpWords == 4) { const uint4 v = *reinterpret_cast<const uint4*>(src); dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
                else {
#pragma unroll
                    for (int i = 0; i < kSpWords; ++i) dst[i] = reinterpret_cast<const uint32_t*>(src)[i];
                }
            } else {
                const uint8_t* srow = a.Sp + t * a.Np + a.n_lo + c0;
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) dst[c >> 2] |= (uint32_t)srow[c] << ((c & 3) * 8);
            }
        };
        load_spikes(0, sb_next);
        unsigned badmask = 0;                            // exp: columns of this warp whose activation left the FP32-safe range

        for (int it = 0; it < ntl; ++it) {
            const int b = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            const int64_t t = (first + (int64_t)it * step) * kTileT + row;
            const float lv = t < a.T ? 1.0f : 0.0f;
#pragma unroll
            for (int i = 0; i < kSpWords; ++i) sb[i] = sb_next[i];
            load_spikes(it + 1, sb_next);
            const bool tr = a.trace && blockIdx.x == 0 && warp == kFirstEpiWarp + (a.debug >> 8) && lane == 0 && it < 32;
            if (tr) a.trace[(2 * 32 + it) * 4 + 0] = clock64();
            mbar_wait(bar_fwd_full, it & 1);
            if (tr) a.trace[(2 * 32 + it) * 4 + 1] = clock64();
            tc_fence_after();
            float d0[kColsPerWarp], d1[kColsPerWarp], d2[kColsPerWarp];
            tmem_ld<kColsPerWarp>(t_lane + 0, d0);
            tmem_ld<kColsPerWarp>(t_lane + 32, d1);
            tmem_ld<kColsPerWarp>(t_lane + 64, d2);
            tmem_ld_wait();
            if (tr) a.trace[1408 + it * 4 + 0] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_fwd_empty);

            // act, Poisson term, residual (padded columns have M = 0, bias = 0 and are ignored later)
            float xs[kColsPerWarp];
            float xmin = 3.0e38f;
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) {
                const float2 cp = cpar[c0 + c];
                xs[c] = fmaf(fmaf(d1[c] + d2[c], 1.0f / kLoScale, d0[c]), cp.x, cp.y);
                if (c0 + c < a.ncols) xmin = fminf(xmin, xs[c]);
                if (NLIN == PYGLM_B200_NLIN_EXP && c0 + c < a.ncols && lv != 0.f && !(xs[c] <= kExpSafe)) badmask |= 1u << c;
            }
            const float dtl = a.dt * lv;                 // bins past the end of the recording contribute nothing
            if (NLIN == PYGLM_B200_NLIN_SOFTPLUS && __all_sync(0xffffffffu, xmin > 17.5f)) {
                // every activation of this warp's block is in the regime where log(1+e^x) rounds to x and
                // sigmoid(x) to 1 in FP32 (e^-17.5 < 2^-25): straight-line code, no divergence, two MUFU per bin
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) {
                    float r = 0.f;
                    if (c0 + c < a.ncols) {              // warp-uniform: padded columns cost nothing
                        const float x = xs[c];
                        const float sv = (float)((sb[c >> 2] >> ((c & 3) * 8)) & 0xffu);     // 0 past the end of the recording
                        float lg, rc;
                        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(x));
                        pll[c] += fmaf(sv * 0.69314718f, lg, -dtl * x);                        // -dt*lam + s*log(lam)
                        r = fmaf(sv, rc, -dtl);                                                 // (s/lam - dt) * 1
                        pgb[c] += r;
                    }
                    d1[c] = r;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) {
                    float term = 0.f, r = 0.f;
                    if (c0 + c < a.ncols) {
                        poisson_terms<NLIN>(xs[c], (float)((sb[c >> 2] >> ((c & 3) * 8)) & 0xffu), a.dt, term, r);
                        r *= lv;
                        pll[c] = fmaf(lv, term, pll[c]);
                        pgb[c] += r;
                    }
                    d1[c] = r;
                }
            }
            // residual planes back to smem as the gradient MMA's B operand (MN-major, 64B swizzle)
            if (tr) a.trace[1408 + it * 4 + 1] = clock64();
            mbar_wait(&bar_r_free[b], ph ^ 1);
            if (tr) a.trace[1408 + it * 4 + 2] = clock64();
            {
                unsigned char* p1 = sR + b * 2 * kRBytes + row * 64;
                unsigned char* p2 = p1 + kRBytes;
#pragma unroll
                for (int u = 0; u < kColsPerWarp / 8; ++u) {
                    uint32_t h1[4], h2[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float ra = d1[8 * u + 2 * k] * kRScale, rb = d1[8 * u + 2 * k + 1] * kRScale;
                        const __half2 hi = __floats2half2_rn(ra, rb);
                        const float2 back = __half22float2(hi);
                        const __half2 lo = __floats2half2_rn((ra - back.x) * kLoScale, (rb - back.y) * kLoScale);
                        h1[k] = *reinterpret_cast<const uint32_t*>(&hi);
                        h2[k] = *reinterpret_cast<const uint32_t*>(&lo);
                    }
                    const int unit = (c0 >> 3) + u;          // 16-byte unit (8 columns) within the 64-byte row
                    *reinterpret_cast<uint4*>(p1 + ((unit ^ sw) << 4)) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                    *reinterpret_cast<uint4*>(p2 + ((unit ^ sw) << 4)) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_r_ready[b]);
            if (tr) a.trace[(2 * 32 + it) * 4 + 2] = clock64();
            if (a.trace && blockIdx.x == 0 && lane == 0 && it < 32) a.trace[384 + warp * 32 + it] = clock64();

            // column sums -> FP64 every F tiles; the previous tile's gradient block -> FP64 (its MMA ran
            // while this tile's epilogue was busy)
            if (((it + 1) % F) == 0 || it == ntl - 1) {
                ll_acc += (double)warp_column_sums<kColsPerWarp>(pll, lane);
                gb_acc += (double)warp_column_sums<kColsPerWarp>(pgb, lane);
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) { pll[c] = 0.f; pgb[c] = 0.f; }
            }
            if (it >= 1) fold_gradient(it - 1);
            if (tr) a.trace[(2 * 32 + it) * 4 + 3] = clock64();
        }

        if (ntl > 0) fold_gradient(ntl - 1);
        if (NLIN == PYGLM_B200_NLIN_EXP && a.flags) {
            badmask = __reduce_or_sync(0xffffffffu, badmask);
            if (lane < kColsPerWarp && ((badmask >> lane) & 1u)) a.flags[a.n_lo + c0 + lane] = 1u;
        }
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) gp[(int64_t)row * kNcol + c0 + c] = gacc[c];
        if (a.nmt > 1) {
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) gp[((int64_t)128 + row) * kNcol + c0 + c] = gacc1[c];
        }
        // ---- per-CTA ll / g_bias partials
        double* sred = reinterpret_cast<double*>(sR);    // residual planes are idle now: [4 quarters][2][32]
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // every epilogue warp is past its last wait
        if (lane < kColsPerWarp) {
            sred[(q * 2 + 0) * kNcol + c0 + lane] = ll_acc;
            sred[(q * 2 + 1) * kNcol + c0 + lane] = gb_acc;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (warp == kFirstEpiWarp) {
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
    if (warp == kProducerWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemAlloc) : "memory");
    }
}

// =============================================================================================
// V2 of the fused kernel for 97..160 features (the headline config C2: 135).  Same algebra, same arithmetic; what
// changes is how much of X is in flight and how long a tile's dependency chain is:
//   * X lives in shared memory as single-PLANE slots (one FP16 plane of one 128-bin tile: two 64-feature blocks in
//     128-byte-swizzled rows + a 16- or 32-feature tail in 32- / 64-byte rows), in a ring of as many slots as fit
//     (5 at 129..144 features): 2.5 tiles resident instead of 2 stages, and the next tile's high plane is already
//     landing while the oldest tile's gradient MMAs still read theirs;
//   * the two 2^-11-scaled products share ONE accumulator (x1 m2 and x2 m1 are issued into the same TMEM columns,
//     likewise x1^T r2 and x2^T r1): 64 instead of 96 columns per block, a third fewer tcgen05.ld in the epilogue
//     and in the gradient fold;
//   * the residual planes form one 128-byte-swizzled operand [r1 | r2] per bin (a full shared-memory wavefront per
//     fetch; the old 64-byte layout left them half empty), single-buffered when that buys another slot;
//   * the gradient block of the tail features (replicated into every TMEM lane group by an LBO = 0 descriptor) is
//     folded by ALL warps every tile, one or two columns each, instead of by one lane quarter every fourth tile --
//     no warp trails the others into the residual barrier any more.
// =============================================================================================
#ifndef PYGLM_TC_SCHED
#define PYGLM_TC_SCHED 0                     // 1: clustered TMEM reads before the residual signal (see the epilogue loop)
#endif
#ifndef PYGLM_TC_TRACE_BUILD
#define PYGLM_TC_TRACE_BUILD 0               // 1: keep the clock64 pipeline stamps (PYGLM_TC_TRACE) in the V2 kernel
#endif
constexpr uint32_t kSw32 = 6;                // UMMA LayoutType::SWIZZLE_32B
#ifndef PYGLM_TC_GBASE
#define PYGLM_TC_GBASE 64
#endif
constexpr int kV2FwdCols = PYGLM_TC_GBASE;   // [hi | lo] activation accumulators in columns 0..63; gradient buffers start here
constexpr int kV2GradTile = 64;              // [hi | lo] gradient accumulators of one 128-feature tile
constexpr int kV2GradBuf = 2 * kV2GradTile;  // tile 0 (features 0..127) + tile 1 (tail)
constexpr int kV2MaxSlots = 6;
#ifndef PYGLM_TC_TAIL_M64
#define PYGLM_TC_TAIL_M64 1                  // M = 64 MMAs for the replicated 16-feature tail block: half the operand fetch (measured 0.154 -> 0.147 ms);
                                             // their accumulator rows sit in lanes 0..15 of every TMEM lane quarter
#endif

struct TcV2Args {
    TcKernelArgs k;
    int nslots;                    // plane slots in the ring
    int rbufs;                     // residual buffers (1 or 2)
};

// TILE = bins per tile: 128, or 112 -- the forward MMA still spans 128 rows (UMMA M), its last 16 rows then read
// whatever follows the block in shared memory (finite FP16 data) and are discarded; what 112 buys is a sixth slot
// (three whole tiles resident) within the 227 KB, so that the load of a tile no longer waits for the gradient MMAs
// of the tile two before it.
template <int TAIL, int TILE> struct V2Geom {
    static constexpr int kBlock = TILE * 128;                        // one 64-feature block of one plane: TILE rows x 128 B
    static constexpr int kTailRowB = 2 * TAIL;                       // bytes per tail row (32 or 64)
    static constexpr int kTailBytes = TILE * kTailRowB;              // X tail block of one plane
    static constexpr int kSlotBytes = (2 * kBlock + kTailBytes + 1023) / 1024 * 1024;   // one plane of one tile
    static constexpr int kMBlock = 64 * 128;                         // [M1 | M2] rows of one 64-feature block
    static constexpr int kMBytes = 2 * kMBlock + 64 * kTailRowB;
    static constexpr int kRBuf = TILE * 128;                         // [r1 | r2]: TILE bins x 128 B
    static constexpr int kUnits = TAIL ? 3 : 2;
    static constexpr int kKSteps = TILE / 16;                        // gradient MMAs step through 16 bins
};

template <int TAIL, int TILE>
__host__ __device__ inline int v2_smem_bytes(int nslots, int rbufs)
{
    using G = V2Geom<TAIL, TILE>;
    return nslots * G::kSlotBytes + G::kMBytes + rbufs * G::kRBuf + 1024;
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float* v)
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}

template <int NLIN, int TAIL, int TILE>
__global__ void __launch_bounds__(kThreads, 1)
tc_fused2_kernel(const __grid_constant__ CUtensorMap tmapW1, const __grid_constant__ CUtensorMap tmapW2,
                 const __grid_constant__ CUtensorMap tmapT1, const __grid_constant__ CUtensorMap tmapT2, TcV2Args va)
{
    using G = V2Geom<TAIL, TILE>;
    constexpr int kBlockBytes = G::kBlock;
    const TcKernelArgs& a = va.k;
#if PYGLM_TC_TRACE_BUILD
    long long* const V2TRACE = a.trace;
#else
    constexpr long long* V2TRACE = nullptr;      // the clock64 stamps are compiled out of the production build
#endif
    const int NS = va.nslots, RB = va.rbufs;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sM = smem + NS * G::kSlotBytes;        // [block 0: 64 rows x 128 B][block 1][tail: 64 rows x kTailRowB]
    unsigned char* sR = sM + G::kMBytes;                   // RB buffers of [128 bins][r1 (32 cols) | r2 (32 cols)]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + RB * G::kRBuf);
    uint64_t* bar_full = bars;                             // [kV2MaxSlots][3]  unit of a slot landed
    uint64_t* bar_empty = bars + 3 * kV2MaxSlots;          // [kV2MaxSlots]     slot consumed by the gradient MMAs
    uint64_t* bar_fwd_full = bar_empty + kV2MaxSlots;      // activation accumulators ready
    uint64_t* bar_fwd_empty = bar_fwd_full + 1;            // ... drained by the epilogue
    uint64_t* bar_r_ready = bar_fwd_empty + 1;             // [2] residual operand written
    uint64_t* bar_r_free = bar_r_ready + 2;                // [2] ... consumed by the gradient MMAs
    uint64_t* bar_g_full = bar_r_free + 2;                 // [2] gradient accumulators of a tile complete
    uint64_t* bar_g_empty = bar_g_full + 2;                // [2] ... folded into FP64
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_g_empty + 2);
    float* sPar = reinterpret_cast<float*>(tmem_slot + 2); // [32] x (1/sm, bias)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 3 * kV2MaxSlots; ++i) mbar_init(&bar_full[i], 1);
        for (int i = 0; i < kV2MaxSlots; ++i) mbar_init(&bar_empty[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_g_full[i], 1); mbar_init(&bar_g_empty[i], kEpiWarps);
            mbar_init(&bar_r_ready[i], kEpiWarps); mbar_init(&bar_r_free[i], 1);
        }
        mbar_init(bar_fwd_full, 1); mbar_init(bar_fwd_empty, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemAlloc) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (TILE != kTileT) {          // rows TILE..127 of a block are read (and discarded) by the forward MMA: keep them finite
        for (int u = threadIdx.x; u < NS * G::kSlotBytes / 16; u += kThreads)
            reinterpret_cast<uint4*>(smem)[u] = make_uint4(0u, 0u, 0u, 0u);
    }
    // weight planes -> shared memory, swizzled by hand: per 64-feature block 64 rows [M1 rows | M2 rows] of 128 B
    // (SWIZZLE_128B), then the tail block with rows of kTailRowB bytes (SWIZZLE_64B / SWIZZLE_32B)
    {
        constexpr int kUnitsPerRow = (128 + TAIL) / 8;                // 16-byte units (8 features) per weight row
        for (int u = threadIdx.x; u < 64 * kUnitsPerRow; u += kThreads) {
            const int r = u / kUnitsPerRow, f8 = u - r * kUnitsPerRow;            // r = plane * 32 + column
            const uint4 val = *reinterpret_cast<const uint4*>(a.Mp + ((int64_t)r * a.Kp + f8 * 8));
            unsigned char* dst;
            if (f8 < 16) {
                const int blk = f8 >> 3, q8 = f8 & 7;
                dst = sM + blk * G::kMBlock + r * 128 + ((q8 ^ (r & 7)) << 4);
            } else if (TAIL == 32) {
                const int q = f8 - 16;
                dst = sM + 2 * G::kMBlock + r * 64 + ((q ^ ((r >> 1) & 3)) << 4);
            } else {
                const int q = f8 - 16;
                dst = sM + 2 * G::kMBlock + r * 32 + ((q ^ ((r >> 2) & 1)) << 4);
            }
            *reinterpret_cast<uint4*>(dst) = val;
        }
        fence_proxy_async();
        if (threadIdx.x < kNcol) {
            sPar[2 * threadIdx.x] = a.colpar[threadIdx.x];
            sPar[2 * threadIdx.x + 1] = a.colpar[kNcol + threadIdx.x];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t first = blockIdx.x, step = gridDim.x;
    const int ntl = first < a.ntiles ? (int)((a.ntiles - first + step - 1) / step) : 0;   // tiles of this CTA
    const int F = a.flush;
    // fill f (0, 1, 2, ...) = plane f & 1 of tile f >> 1, lands in slot f % NS; its barriers' parity is (f / NS) & 1

    if (warp == kProducerWarp) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int slot = 0; uint32_t par = 1;              // parity on which the slot's empty barrier is waited
            for (int it = 0; it < ntl; ++it) {
                const int row0 = (int)((first + (int64_t)it * step) * TILE);
                for (int plane = 0; plane < 2; ++plane) {
                    mbar_wait_relaxed(&bar_empty[slot], par, a.producer_sleep_ns);
                    if (plane == 0 && V2TRACE && blockIdx.x == 0 && it < 32) a.trace[(0 * 32 + it) * 4 + 0] = clock64();
                    unsigned char* st = smem + slot * G::kSlotBytes;
                    const CUtensorMap* mw = plane ? &tmapW2 : &tmapW1;
                    for (int u = 0; u < 2; ++u) {
                        uint64_t* fb = &bar_full[slot * 3 + u];
                        mbar_arrive_expect_tx(fb, kBlockBytes);
                        tma_load_2d(st + u * kBlockBytes, mw, fb, u * 64, row0);
                    }
                    if (TAIL) {
                        uint64_t* fb = &bar_full[slot * 3 + 2];
                        mbar_arrive_expect_tx(fb, G::kTailBytes);
                        tma_load_2d(st + 2 * kBlockBytes, plane ? &tmapT2 : &tmapT1, fb, 128, row0);
                    }
                    if (++slot == NS) { slot = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == kFwdWarp) {
        // ================================ forward MMA issuer ==========================
        if (lane == 0) {
            constexpr uint32_t idesc_f64 = umma_idesc(128, 64, 0, 0);     // A: X K-major,  B: [M1|M2] K-major
            constexpr uint32_t idesc_f32 = umma_idesc(128, 32, 0, 0);
            const uint32_t sMb = smem_u32(sM);
            const uint32_t t_f = tmem_base;
            int slot = 0; uint32_t par = 0;
            for (int it = 0; it < ntl; ++it) {
                mbar_wait_role(bar_fwd_empty, (it & 1) ^ 1);
                for (int plane = 0; plane < 2; ++plane) {
                    const uint32_t sX = smem_u32(smem + slot * G::kSlotBytes);
                    // plane 0: x1 [m1 | m2] -> columns 0..63 (first MMA overwrites); plane 1: x2 m1 -> added to columns 32..63
                    const uint32_t t_d = plane ? t_f + 32 : t_f;
                    const uint32_t idesc = plane ? idesc_f32 : idesc_f64;
                    for (int u = 0; u < 2; ++u) {
                        mbar_wait_role(&bar_full[slot * 3 + u], par);
                        if (plane == 0 && u == 0 && V2TRACE && blockIdx.x == 0 && it < 32) a.trace[(1 * 32 + it) * 4 + 0] = clock64();
                        tc_fence_after();
                        const uint64_t wx = umma_desc_sw128(sX + u * kBlockBytes);
                        const uint64_t wm = umma_desc_sw128(sMb + u * G::kMBlock);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16(t_d, wx + 2 * ks, wm + 2 * ks, idesc, (plane | u | ks) ? 1u : 0u);
                    }
                    if (TAIL) {
                        mbar_wait_role(&bar_full[slot * 3 + 2], par);
                        tc_fence_after();
                        if (TAIL == 32) {
                            const uint64_t dx = umma_desc(sX + 2 * kBlockBytes, 16, 512), dm = umma_desc(sMb + 2 * G::kMBlock, 16, 512);
                            umma_f16(t_d, dx, dm, idesc, 1u);
                            umma_f16(t_d, dx + 2, dm + 2, idesc, 1u);
                        } else {
                            const uint64_t dx = umma_desc_layout(sX + 2 * kBlockBytes, 16, 256, kSw32);
                            const uint64_t dm = umma_desc_layout(sMb + 2 * G::kMBlock, 16, 256, kSw32);
                            umma_f16(t_d, dx, dm, idesc, 1u);
                        }
                    }
                    if (++slot == NS) { slot = 0; par ^= 1; }
                }
                umma_commit(bar_fwd_full);
                if (V2TRACE && blockIdx.x == 0 && it < 32) a.trace[(1 * 32 + it) * 4 + 1] = clock64();
            }
        }
    } else if (warp == kBwdWarp) {
        // ================================ gradient MMA issuer =========================
        if (lane == 0) {
            constexpr uint32_t idesc_b64 = umma_idesc(128, 64, 1, 1);     // A: X MN-major, B: [r1|r2] MN-major
            constexpr uint32_t idesc_b32 = umma_idesc(128, 32, 1, 1);
            int slot = 0; uint32_t par = 0;
            int b = 0; uint32_t rph = 0;                                               // b = j % RB, rph = (j / RB) & 1, kept as running counters
            for (int j = 0; j < ntl; ++j, rph ^= (b + 1 == RB) ? 1u : 0u, b = (b + 1 == RB) ? 0 : b + 1) {
                const int gb = j & 1;
                const int sA = slot;
                const int sB = (slot + 1 == NS) ? 0 : slot + 1;
                const uint32_t parB = (slot + 1 == NS) ? par ^ 1 : par;
                for (int u = 0; u < G::kUnits; ++u) {                                  // landed long ago
                    mbar_wait_role(&bar_full[sA * 3 + u], par);
                    mbar_wait_role(&bar_full[sB * 3 + u], parB);
                }
                mbar_wait_role(&bar_r_ready[b], rph);
                mbar_wait_role(&bar_g_empty[gb], ((j >> 1) & 1) ^ 1);
                if (V2TRACE && blockIdx.x == 0 && j < 32) a.trace[(1 * 32 + j) * 4 + 2] = clock64();
                tc_fence_after();
                const uint32_t sX1 = smem_u32(smem + sA * G::kSlotBytes), sX2 = smem_u32(smem + sB * G::kSlotBytes);
                const uint64_t dr = umma_desc_layout(smem_u32(sR + b * G::kRBuf), 16, 1024, kSw128);    // one 64-wide atom
                const uint32_t t_g = tmem_base + kV2FwdCols + gb * kV2GradBuf;
                {   // tile 0: features 0..127 = the two 64-feature blocks (LBO steps from one to the other)
                    const uint64_t d1 = umma_desc_layout(sX1, kBlockBytes, 1024, kSw128);
                    const uint64_t d2 = umma_desc_layout(sX2, kBlockBytes, 1024, kSw128);
#pragma unroll
                    for (int ks = 0; ks < G::kKSteps; ++ks) {                         // 16 bins per step: 2048 B of both operands
                        const uint64_t off = (uint64_t)(ks * 2048 >> 4);
                        umma_f16(t_g, d1 + off, dr + off, idesc_b64, ks ? 1u : 0u);   // X1^T [r1 | r2] -> hi, lo
                        umma_f16(t_g + 32, d2 + off, dr + off, idesc_b32, 1u);        // X2^T r1 -> lo
                    }
                }
                if (TAIL) {
                    // tail features: LBO = 0 makes every group of TAIL TMEM lanes a copy of the same TAIL gradient rows, so
                    // the epilogue warps of all four lane quarters can each fold a share of them
                    const uint32_t kstep = TAIL == 32 ? 1024u : 512u;                 // sixteen bins of the tail block
                    const uint64_t d1 = TAIL == 32 ? umma_desc(sX1 + 2 * kBlockBytes, 0, 512)
                                                   : umma_desc_layout(sX1 + 2 * kBlockBytes, 0, 256, kSw32);
                    const uint64_t d2 = TAIL == 32 ? umma_desc(sX2 + 2 * kBlockBytes, 0, 512)
                                                   : umma_desc_layout(sX2 + 2 * kBlockBytes, 0, 256, kSw32);
                    constexpr uint32_t idesc_t64 = umma_idesc((PYGLM_TC_TAIL_M64 && TAIL == 16) ? 64 : 128, 64, 1, 1);
                    constexpr uint32_t idesc_t32 = umma_idesc((PYGLM_TC_TAIL_M64 && TAIL == 16) ? 64 : 128, 32, 1, 1);
#pragma unroll
                    for (int ks = 0; ks < G::kKSteps; ++ks) {
                        const uint64_t offx = (uint64_t)(ks * kstep >> 4), off = (uint64_t)(ks * 2048 >> 4);
                        umma_f16(t_g + kV2GradTile, d1 + offx, dr + off, idesc_t64, ks ? 1u : 0u);
                        umma_f16(t_g + kV2GradTile + 32, d2 + offx, dr + off, idesc_t32, 1u);
                    }
                }
                umma_commit(&bar_empty[sA]);
                umma_commit(&bar_empty[sB]);
                umma_commit(&bar_r_free[b]);
                umma_commit(&bar_g_full[gb]);
                if (V2TRACE && blockIdx.x == 0 && j < 32) a.trace[(1 * 32 + j) * 4 + 3] = clock64();
                slot += 2;
                if (slot >= NS) { slot -= NS; par ^= 1; }
            }
        }
    } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kEpiWarps) {
        // ================================ epilogue warps ==============================
        const int q = warp & 3;
        const int cg = (warp - kFirstEpiWarp) >> 2;
        const int c0 = cg * kColsPerWarp;                // first column of this warp
        const int row = q * 32 + lane;                   // row of the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + c0;
        const float2* cpar = reinterpret_cast<const float2*>(sPar);     // (1/sm, bias) per column
        const int64_t per_cta = (int64_t)(128 + TAIL) * kNcol + 2 * kNcol;
        double* gp = a.part + (int64_t)blockIdx.x * per_cta;
        double gacc[kColsPerWarp];                       // gradient tile 0: this thread's feature row x its columns
        double gtail[2] = {0.0, 0.0};                    // tail: columns c0 + 2q, c0 + 2q + 1 (TAIL 16: one of them)
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) gacc[c] = 0.0;
        static_assert(kColsPerWarp == 8, "the tail fold assigns two of eight columns to each lane quarter");
        // The gradient block of tile j is read out of TMEM (fold_load) and added into FP64 registers (fold_add).  A
        // tcgen05.ld issued while MMAs are in flight waits behind them (measured: ~2k cycles when it lands in a gradient
        // burst), so where it sits in the loop matters; after the residual signal was the best position measured.
        float g0[kColsPerWarp], g1[kColsPerWarp], h0[2], h1[2];
        auto fold_load = [&](int j) {
            const int gb = j & 1;
            const uint32_t t_g = t_lane + kV2FwdCols + gb * kV2GradBuf;
            mbar_wait(&bar_g_full[gb], (j >> 1) & 1);
            tc_fence_after();
            tmem_ld<kColsPerWarp>(t_g + 0, g0);
            tmem_ld<kColsPerWarp>(t_g + 32, g1);
            if (TAIL) {
                tmem_ld2(t_g + kV2GradTile + 2 * q, h0);
                tmem_ld2(t_g + kV2GradTile + 32 + 2 * q, h1);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_g_empty[gb]);
        };
        auto fold_add = [&]() {
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) gacc[c] += (double)fmaf(g1[c], 1.0f / kLoScale, g0[c]);
            if (TAIL == 32 || (PYGLM_TC_TAIL_M64 && TAIL == 16)) {
                gtail[0] += (double)fmaf(h1[0], 1.0f / kLoScale, h0[0]);
                gtail[1] += (double)fmaf(h1[1], 1.0f / kLoScale, h0[1]);
            } else if (TAIL == 16) {
                const bool up = lane >= 16;              // lanes 16..31 hold a second copy of the 16 rows: they take the odd column
                gtail[0] += (double)fmaf(up ? h1[1] : h1[0], 1.0f / kLoScale, up ? h0[1] : h0[0]);
            }
        };
        double ll_acc = 0.0, gb_acc = 0.0;               // lane l accumulates column c0 + (l % kColsPerWarp)
        float pll[kColsPerWarp], pgb[kColsPerWarp];
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) { pll[c] = 0.f; pgb[c] = 0.f; }

        constexpr int kSpWords = kColsPerWarp / 4;
        uint32_t sb[kSpWords], sb_next[kSpWords];
        const bool vec_ok = ((a.n_lo + c0) % kColsPerWarp) == 0;
        auto load_spikes = [&](int it, uint32_t (&dst)[kSpWords]) {
            const int64_t t = (first + (int64_t)it * step) * TILE + row;
            const bool ok = it < ntl && t < a.T && row < TILE;
#pragma unroll
            for (int i = 0; i < kSpWords; ++i) dst[i] = 0;
            if (!ok) return;
            const uint8_t* src = a.Sp + t * a.Np + a.n_lo + c0;
            if (vec_ok) {
                const uint2 v = *reinterpret_cast<const uint2*>(src);
                dst[0] = v.x; dst[1] = v.y;
            } else {
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) dst[c >> 2] |= (uint32_t)src[c] << ((c & 3) * 8);
            }
        };
        load_spikes(0, sb_next);
        unsigned badmask = 0;

#if PYGLM_TC_TRACE_BUILD
#define PYGLM_WSTAMP(K) do { if (V2TRACE && blockIdx.x == 0 && lane == 0 && it < 24) a.trace[2048 + ((warp - kFirstEpiWarp) * 24 + it) * 8 + (K)] = clock64(); } while (0)
#else
#define PYGLM_WSTAMP(K) do { } while (0)
#endif
#if PYGLM_TC_SCHED
        // Clustered TMEM reads: every tcgen05.ld of an iteration -- the gradient block of the previous tile AND the
        // activation accumulators of the NEXT tile -- is issued at one point, just before this warp signals its residual
        // rows.  At that point the tensor pipe is idle by construction (the previous gradient MMAs and the next forward
        // MMAs are complete, the following ones cannot be issued before these loads are done), so no load ever queues
        // behind a burst of MMAs; after the signal the gradient MMAs of this tile and the forward MMAs of the tile after
        // next run back to back while the epilogue warps do the next tile's math.
        float d0n[kColsPerWarp], d1n[kColsPerWarp];
        if (ntl > 0) {
            mbar_wait(bar_fwd_full, 0);
            tc_fence_after();
            tmem_ld<kColsPerWarp>(t_lane + 0, d0n);
            tmem_ld<kColsPerWarp>(t_lane + 32, d1n);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_fwd_empty);
        }
#endif
        // b = it % RB, rph = (it / RB) & 1 and the flush period as running counters: an integer division by a kernel argument
        // costs ~30 instructions, and this loop is instruction-issue-bound
        int b = 0, fcnt = 0; uint32_t rph = 0;
        for (int it = 0; it < ntl; ++it, rph ^= (b + 1 == RB) ? 1u : 0u, b = (b + 1 == RB) ? 0 : b + 1) {
            PYGLM_WSTAMP(0);
            const int64_t t = (first + (int64_t)it * step) * TILE + row;
            const float lv = (t < a.T && row < TILE) ? 1.0f : 0.0f;
#pragma unroll
            for (int i = 0; i < kSpWords; ++i) sb[i] = sb_next[i];
            load_spikes(it + 1, sb_next);
            const bool tr = V2TRACE && blockIdx.x == 0 && warp == kFirstEpiWarp + (a.debug >> 8) && lane == 0 && it < 32;
            if (tr) a.trace[(2 * 32 + it) * 4 + 0] = clock64();
            float d0[kColsPerWarp], d1[kColsPerWarp];
#if PYGLM_TC_SCHED
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) { d0[c] = d0n[c]; d1[c] = d1n[c]; }
            PYGLM_WSTAMP(1);
#else
            mbar_wait(bar_fwd_full, it & 1);
            if (tr) a.trace[(2 * 32 + it) * 4 + 1] = clock64();
            PYGLM_WSTAMP(1);
            tc_fence_after();
            tmem_ld<kColsPerWarp>(t_lane + 0, d0);
            tmem_ld<kColsPerWarp>(t_lane + 32, d1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_fwd_empty);
#endif
            PYGLM_WSTAMP(2);

            float xs[kColsPerWarp];
            float xmin = 3.0e38f;
#pragma unroll
            for (int c = 0; c < kColsPerWarp; ++c) {
                const float2 cp = cpar[c0 + c];
                xs[c] = fmaf(fmaf(d1[c], 1.0f / kLoScale, d0[c]), cp.x, cp.y);
                if (TILE != kTileT && row >= TILE) xs[c] = 30.0f;      // discarded rows: any finite activation
                if (c0 + c < a.ncols) xmin = fminf(xmin, xs[c]);
                if (NLIN == PYGLM_B200_NLIN_EXP && c0 + c < a.ncols && lv != 0.f && !(xs[c] <= kExpSafe)) badmask |= 1u << c;
            }
            const float dtl = a.dt * lv;                 // bins past the end of the recording contribute nothing
            if (NLIN == PYGLM_B200_NLIN_SOFTPLUS && __all_sync(0xffffffffu, xmin > 17.5f)) {
                // One vote per warp and tile: does any of its 32 bins x 8 columns hold a spike?  (A vote per column skips the
                // spike math of the half of the column slices that hold none, but its eight votes and reconvergence points
                // cost more issue slots than they save: 0.1454 -> 0.1403 ms, A/B on one box.)
                const bool anysp = __any_sync(0xffffffffu, (sb[0] | sb[kSpWords - 1]) != 0u);
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) {
                    float r = 0.f;
                    if (c0 + c < a.ncols) {              // warp-uniform: padded columns cost nothing
                        const float x = xs[c];
                        const unsigned sbyte = (sb[c >> 2] >> ((c & 3) * 8)) & 0xffu;          // 0 past the end of the recording
                        pll[c] = fmaf(-dtl, x, pll[c]);                                         // -dt*lam
                        r = -dtl;                                                               // -dt * f', f' = 1
                        if (anysp) {
                            const float sv = (float)sbyte;
                            float lg, rc;
                            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(x));
                            pll[c] = fmaf(sv * 0.69314718f, lg, pll[c]);                        // + s*log(lam)
                            r = fmaf(sv, rc, r);                                                // + s/lam
                        }
                        pgb[c] += r;
                    }
                    d1[c] = r;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) {
                    float term = 0.f, r = 0.f;
                    if (c0 + c < a.ncols) {
                        poisson_terms<NLIN>(xs[c], (float)((sb[c >> 2] >> ((c & 3) * 8)) & 0xffu), a.dt, term, r);
                        r *= lv;
                        pll[c] = fmaf(lv, term, pll[c]);
                        pgb[c] += r;
                    }
                    d1[c] = r;
                }
            }
            // residual operand [r1 | r2] of this bin: one 128-byte row, 16-byte units 0..3 = r1 (8 columns each), 4..7 = r2,
            // unit index XOR (row & 7) (SWIZZLE_128B); this warp owns unit cg of each half
            if (tr) a.trace[1408 + it * 4 + 1] = clock64();
            PYGLM_WSTAMP(3);
#if PYGLM_TC_SCHED
            {
                const bool hasg = it >= 1, hasf = it + 1 < ntl;
                const int gbp = (it - 1) & 1;
                if (hasg) {
                    const uint32_t t_g = t_lane + kV2FwdCols + gbp * kV2GradBuf;
                    mbar_wait(&bar_g_full[gbp], ((it - 1) >> 1) & 1);
                    tc_fence_after();
                    tmem_ld<kColsPerWarp>(t_g + 0, g0);
                    tmem_ld<kColsPerWarp>(t_g + 32, g1);
                    if (TAIL) {
                        tmem_ld2(t_g + kV2GradTile + 2 * q, h0);
                        tmem_ld2(t_g + kV2GradTile + 32 + 2 * q, h1);
                    }
                }
                if (hasf) {
                    mbar_wait(bar_fwd_full, (it + 1) & 1);
                    tc_fence_after();
                    tmem_ld<kColsPerWarp>(t_lane + 0, d0n);
                    tmem_ld<kColsPerWarp>(t_lane + 32, d1n);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (hasg) mbar_arrive(&bar_g_empty[gbp]);
                    if (hasf) mbar_arrive(bar_fwd_empty);
                }
            }
#endif
            mbar_wait(&bar_r_free[b], rph ^ 1);
            PYGLM_WSTAMP(4);
            if (tr) a.trace[1408 + it * 4 + 2] = clock64();
            if (TILE == kTileT || row < TILE) {
                unsigned char* prow = sR + b * G::kRBuf + row * 128;
                uint32_t h1[4], h2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float ra = d1[2 * k] * kRScale, rb = d1[2 * k + 1] * kRScale;
                    const __half2 hi = __floats2half2_rn(ra, rb);
                    const float2 back = __half22float2(hi);
                    const __half2 lo = __floats2half2_rn((ra - back.x) * kLoScale, (rb - back.y) * kLoScale);
                    h1[k] = *reinterpret_cast<const uint32_t*>(&hi);
                    h2[k] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                const int sw = row & 7;
                *reinterpret_cast<uint4*>(prow + ((cg ^ sw) << 4)) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                *reinterpret_cast<uint4*>(prow + (((4 + cg) ^ sw) << 4)) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_r_ready[b]);
            PYGLM_WSTAMP(5);
            if (tr) a.trace[(2 * 32 + it) * 4 + 2] = clock64();
            if (V2TRACE && blockIdx.x == 0 && lane == 0 && it < 32) a.trace[384 + warp * 32 + it] = clock64();

            if (++fcnt == F || it == ntl - 1) {
                fcnt = 0;
                ll_acc += (double)warp_column_sums<kColsPerWarp>(pll, lane);
                gb_acc += (double)warp_column_sums<kColsPerWarp>(pgb, lane);
#pragma unroll
                for (int c = 0; c < kColsPerWarp; ++c) { pll[c] = 0.f; pgb[c] = 0.f; }
            }
            PYGLM_WSTAMP(6);
#if PYGLM_TC_SCHED
            if (it >= 1) fold_add();
#else
            if (it >= 1) { fold_load(it - 1); fold_add(); }
#endif
            PYGLM_WSTAMP(7);
            if (tr) a.trace[(2 * 32 + it) * 4 + 3] = clock64();
        }
#undef PYGLM_WSTAMP

        if (ntl > 0) { fold_load(ntl - 1); fold_add(); }
        if (NLIN == PYGLM_B200_NLIN_EXP && a.flags) {
            badmask = __reduce_or_sync(0xffffffffu, badmask);
            if (lane < kColsPerWarp && ((badmask >> lane) & 1u)) a.flags[a.n_lo + c0 + lane] = 1u;
        }
#pragma unroll
        for (int c = 0; c < kColsPerWarp; ++c) gp[(int64_t)row * kNcol + c0 + c] = gacc[c];
        if (TAIL == 16 && PYGLM_TC_TAIL_M64) {           // M = 64: every lane quarter holds the 16 rows once, in its lanes 0..15
            if (lane < 16) {
                gp[((int64_t)128 + lane) * kNcol + c0 + 2 * q] = gtail[0];
                gp[((int64_t)128 + lane) * kNcol + c0 + 2 * q + 1] = gtail[1];
            }
        } else if (TAIL == 32) {
            gp[((int64_t)128 + lane) * kNcol + c0 + 2 * q] = gtail[0];
            gp[((int64_t)128 + lane) * kNcol + c0 + 2 * q + 1] = gtail[1];
        } else if (TAIL == 16) {
            gp[((int64_t)128 + (lane & 15)) * kNcol + c0 + 2 * q + (lane >> 4)] = gtail[0];
        }
        // ---- per-CTA ll / g_bias partials
        double* sred = reinterpret_cast<double*>(sR);    // the residual operand is idle now: [4 quarters][2][32]
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // every epilogue warp is past its last wait
        if (lane < kColsPerWarp) {
            sred[(q * 2 + 0) * kNcol + c0 + lane] = ll_acc;
            sred[(q * 2 + 1) * kNcol + c0 + lane] = gb_acc;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (warp == kFirstEpiWarp) {
            double l = 0.0, g = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { l += sred[(k * 2 + 0) * kNcol + lane]; g += sred[(k * 2 + 1) * kNcol + lane]; }
            double* lp = gp + (int64_t)(128 + TAIL) * kNcol;
            lp[lane] = l;
            lp[kNcol + lane] = g;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kProducerWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemAlloc) : "memory");
    }
}

// V2 partials: per CTA [128 + TAIL feature rows][32], then [32] ll, [32] g_bias; summed in CTA order (deterministic)
__global__ void __launch_bounds__(32 * 32)
tc_final2_kernel(const double* __restrict__ part, int nctas, int nrows, int N, int B, int F, int n_lo, int ncols,
                 const float* __restrict__ sx, const int8_t* __restrict__ A, const double* __restrict__ W,
                 double* __restrict__ out_ll, double* __restrict__ out_gb, double* __restrict__ out_gw)
{
    __shared__ double sh[32][2 * kNcol];
    const int64_t NS = (int64_t)N * B, NB = NS + F;
    const int64_t per_cta = (int64_t)nrows * kNcol + 2 * kNcol;
    const int nl = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const bool tail = blockIdx.x == NB;                      // the ll / g_bias block
    const int64_t off = tail ? (int64_t)nrows * kNcol : (int64_t)blockIdx.x * kNcol;
    double s0 = 0.0, s1 = 0.0;
    for (int c = slice; c < nctas; c += 32) {
        const double* pc = part + c * per_cta + off + nl;
        s0 += pc[0];
        if (tail) s1 += pc[kNcol];
    }
    sh[slice][nl] = s0;
    sh[slice][kNcol + nl] = s1;
    __syncthreads();
    if (slice != 0 || nl >= ncols) return;
#pragma unroll
    for (int k = 1; k < 32; ++k) { s0 += sh[k][nl]; s1 += sh[k][kNcol + nl]; }
    if (tail) {
        out_ll[nl] = s0;
        if (out_gb) out_gb[nl] = s1;                         // residual column sums are carried unscaled
    } else if (out_gw) {
        const int64_t j = blockIdx.x;
        const int n = n_lo + nl, pre = (int)(j / B);
        const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
        const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
        out_gw[(int64_t)nl * NB + j] = (a * ww) * s0 / ((double)sx[j] * (double)kRScale);
    }
}

// tc_final2_kernel with the sum over ranks folded in (time-sharded evaluation, csrc/allreduce.cu's protocol with one
// block per feature row): a block writes its 32 column values into its slot of every peer's receive buffer, raises its
// flag there, waits for the peers' flags of the same row and adds the slots in rank order -- the evaluation's last
// kernel IS the collective, one launch and one flag round trip instead of two launches.  `out` is the contiguous result
// vector [ll (N) | g_bias (N) | g_w (N x NB)]; all N neurons (N <= 32) are evaluated.
__global__ void __launch_bounds__(32 * 32)
tc_final2_allreduce_kernel(const double* __restrict__ part, int nctas, int nrows, int N, int B, int F,
                           const float* __restrict__ sx, const int8_t* __restrict__ A, const double* __restrict__ W,
                           ArEpoch ar, double* __restrict__ out)
{
    __shared__ double sh[32][2 * kNcol];
    const int64_t NS = (int64_t)N * B, NB = NS + F;
    const int64_t per_cta = (int64_t)nrows * kNcol + 2 * kNcol;
    const int nl = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const bool tail = blockIdx.x == NB;                      // the ll / g_bias block
    const int64_t off = tail ? (int64_t)nrows * kNcol : (int64_t)blockIdx.x * kNcol;
    double s0 = 0.0, s1 = 0.0;
    for (int c = slice; c < nctas; c += 32) {
        const double* pc = part + c * per_cta + off + nl;
        s0 += pc[0];
        if (tail) s1 += pc[kNcol];
    }
    sh[slice][nl] = s0;
    sh[slice][kNcol + nl] = s1;
    __syncthreads();
    const bool mine = slice == 0 && nl < N;
    int64_t i0 = 0, i1 = 0;                                  // positions of this thread's value(s) in the result vector
    if (mine) {
#pragma unroll
        for (int k = 1; k < 32; ++k) { s0 += sh[k][nl]; s1 += sh[k][kNcol + nl]; }
        if (tail) {
            i0 = nl; i1 = N + nl;
        } else {
            const int64_t j = blockIdx.x;
            const int pre = (int)(j / B);
            const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + nl] : 1.0;
            const double ww = (W && j < NS) ? W[(int64_t)pre * N + nl] : 1.0;
            s0 = (a * ww) * s0 / ((double)sx[j] * (double)kRScale);
            i0 = 2 * (int64_t)N + (int64_t)nl * NB + j;
        }
        const int64_t slot = ((int64_t)(ar.epoch & 1u) * ar.world + ar.rank) * ar.cap;
        for (int r = 0; r < ar.world; ++r) {
            double* dst = ar.peers.recv[r] + slot;
            dst[i0] = s0;
            if (tail) dst[i1] = s1;
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < ar.world) {
        st_release_sys(ar.peers.flag[threadIdx.x] + (int64_t)ar.rank * kArMaxBlocks + blockIdx.x, ar.epoch);
        const unsigned* fl = ar.peers.flag[ar.rank] + (int64_t)threadIdx.x * kArMaxBlocks + blockIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(fl) - ar.epoch) < 0) {
            if (clock64() - t0 > (1ll << 37)) __trap();       // a peer never arrived: fail loudly, do not hang
        }
    }
    __syncthreads();
    if (mine) {
        const double* src = ar.peers.recv[ar.rank] + (int64_t)(ar.epoch & 1u) * ar.world * ar.cap;
        double t0s = 0.0, t1s = 0.0;
        for (int r = 0; r < ar.world; ++r) {
            t0s += __ldcg(src + (int64_t)r * ar.cap + i0);
            if (tail) t1s += __ldcg(src + (int64_t)r * ar.cap + i1);
        }
        out[i0] = t0s;
        if (tail) out[i1] = t1s;
    }
}

// Sum the per-CTA partials in a fixed order and undo the scales.  One block per feature row j
// (plus one for ll / g_bias): 32 columns x 32 slices of the CTA range, combined slice 0..31.
constexpr int kFinalSlices = 32;
__global__ void __launch_bounds__(32 * kFinalSlices)
tc_final_kernel(const double* __restrict__ part, int nctas, int nmt, int N, int B, int F, int n_lo, int ncols,
                const float* __restrict__ sx, const int8_t* __restrict__ A, const double* __restrict__ W,
                double* __restrict__ out_ll, double* __restrict__ out_gb, double* __restrict__ out_gw)
{
    __shared__ double sh[kFinalSlices][2 * kNcol];
    const int64_t NS = (int64_t)N * B, NB = NS + F;
    const int64_t per_cta = (int64_t)nmt * 128 * kNcol + 2 * kNcol;
    const int nl = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const bool tail = blockIdx.x == NB;                      // the ll / g_bias block
    const int64_t off = tail ? (int64_t)nmt * 128 * kNcol : (int64_t)blockIdx.x * kNcol;
    const bool rep = !tail && blockIdx.x >= 128;             // gradient tile 1: four lane-quarter copies, 32 rows apart
    double s0 = 0.0, s1 = 0.0;
    for (int c = slice; c < nctas; c += kFinalSlices) {
        const double* pc = part + c * per_cta + off + nl;
        s0 += pc[0];
        if (rep) s0 += pc[32 * kNcol] + pc[64 * kNcol] + pc[96 * kNcol];
        if (tail) s1 += pc[kNcol];
    }
    sh[slice][nl] = s0;
    sh[slice][kNcol + nl] = s1;
    __syncthreads();
    if (slice != 0 || nl >= ncols) return;
#pragma unroll
    for (int k = 1; k < kFinalSlices; ++k) { s0 += sh[k][nl]; s1 += sh[k][kNcol + nl]; }
    if (tail) {
        out_ll[nl] = s0;
        if (out_gb) out_gb[nl] = s1;                         // residual column sums are carried unscaled
    } else if (out_gw) {
        const int64_t j = blockIdx.x;
        const int n = n_lo + nl, pre = (int)(j / B);
        const double a = (A && j < NS) ? (double)A[(int64_t)pre * N + n] : 1.0;
        const double ww = (W && j < NS) ? W[(int64_t)pre * N + n] : 1.0;
        out_gw[(int64_t)nl * NB + j] = (a * ww) * s0 / ((double)sx[j] * (double)kRScale);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D FP16 tensor map with 64B swizzle: dim0 fastest (elements), row pitch in elements, box {box0, box1}
int tc_make_map_2d(void* map_out, const void* base, int64_t dim0, int64_t dim1, int64_t pitch_elems, int box0, int box1,
                   int swizzle)          // 0: SWIZZLE_64B, 1: SWIZZLE_128B, 2: SWIZZLE_32B
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
    cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for dims %lld x %lld pitch %lld box %d x %d", (int)r,
                  (long long)dim0, (long long)dim1, (long long)pitch_elems, box0, box1);
        return PYGLM_B200_ECUDA;
    }
    return PYGLM_B200_OK;
}

// (re)encode the twelve tensor maps over the X planes for `rows` valid rows (rows beyond read as zero)
static int tc_encode_maps(TcWorkspace& ws, int64_t rows)
{
    CUtensorMap* maps = static_cast<CUtensorMap*>(ws.tmaps);
    int rc;
    if ((rc = tc_make_map_2d(&maps[0], ws.X1, ws.ldp, rows, ws.ldp, kChunkF, kTileT))) return rc;
    if ((rc = tc_make_map_2d(&maps[1], ws.X2, ws.ldp, rows, ws.ldp, kChunkF, kTileT))) return rc;
    if ((rc = tc_make_map_2d(&maps[2], ws.X1, ws.ldp, rows, ws.ldp, 64, kTileT, true))) return rc;      // wide chunks, 128B swizzle
    if ((rc = tc_make_map_2d(&maps[3], ws.X2, ws.ldp, rows, ws.ldp, 64, kTileT, true))) return rc;
    if ((rc = tc_make_map_2d(&maps[4], ws.X1, ws.ldp, rows, ws.ldp, 16, kTileT, 2))) return rc;         // 16-feature tail, 32B swizzle
    if ((rc = tc_make_map_2d(&maps[5], ws.X2, ws.ldp, rows, ws.ldp, 16, kTileT, 2))) return rc;
    for (int pl = 0; pl < 2; ++pl) {                      // the same three box shapes with 112-bin boxes (fused kernel V2)
        const __half* base = pl ? ws.X2 : ws.X1;
        if ((rc = tc_make_map_2d(&maps[6 + pl], base, ws.ldp, rows, ws.ldp, kChunkF, 112))) return rc;
        if ((rc = tc_make_map_2d(&maps[8 + pl], base, ws.ldp, rows, ws.ldp, 64, 112, 1))) return rc;
        if ((rc = tc_make_map_2d(&maps[10 + pl], base, ws.ldp, rows, ws.ldp, 16, 112, 2))) return rc;
    }
    ws.map_rows = rows;
    return PYGLM_B200_OK;
}

// allocate everything the tensor-core path keeps per dataset (planes of `plane_rows` rows, scales, padded spikes, maps)
static int alloc_planes(TcWorkspace& ws, const uint8_t* S, int64_t T, int N, int halo, int NB, cudaStream_t stream,
                        int64_t plane_rows = -1)
{
    if (plane_rows < 0) plane_rows = T;
    ws.ldp = round_up(NB, 8);
    int dev = 0;
    PYGLM_CUDA(cudaGetDevice(&dev));
    PYGLM_CUDA(cudaDeviceGetAttribute(&ws.num_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t plane = (size_t)plane_rows * ws.ldp;
    note_allocation();
    PYGLM_CUDA(cudaMalloc(&ws.X1, plane * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.X2, plane * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.sx, NB * sizeof(float)));
    PYGLM_CUDA(cudaMalloc(&ws.colmax, (size_t)(NB + N) * sizeof(unsigned)));   // per feature, then per presynaptic column
    PYGLM_CUDA(cudaMalloc(&ws.Mp, (size_t)2 * kNcol * kMaxChunks * kChunkF * sizeof(__half)));
    PYGLM_CUDA(cudaMalloc(&ws.colpar, 2 * kNcol * sizeof(float)));
    ws.Np = (int)round_up(N, 32) + 32;            // slack: a column group may start anywhere below N
    PYGLM_CUDA(cudaMalloc(&ws.colflag, (size_t)N * sizeof(unsigned)));
    PYGLM_CUDA(cudaMemsetAsync(ws.colflag, 0, (size_t)N * sizeof(unsigned), stream));
    PYGLM_CUDA(cudaMalloc(&ws.Sp, (size_t)T * ws.Np));
    PYGLM_CUDA(cudaMemsetAsync(ws.Sp, 0, (size_t)T * ws.Np, stream));
    PYGLM_CUDA(cudaMemcpy2DAsync(ws.Sp, ws.Np, S + (size_t)halo * N, N, N, T, cudaMemcpyDeviceToDevice, stream));
    ws.tmaps = malloc(12 * sizeof(CUtensorMap));
    if (!ws.tmaps) { set_error("host allocation failed"); return PYGLM_B200_ENOMEM; }
    return tc_encode_maps(ws, plane_rows);
}

// per-feature scales: analytic for the N*B spike-history features, from the data for F stimulus features (X resident)
static int tc_compute_scales(TcWorkspace& ws, const uint8_t* S, int64_t srows, int N, const double* d_ibasis, int R, int B,
                             const float* X, int64_t T, int F, int64_t ldx, cudaStream_t stream)
{
    const int NS = N * B, NB = NS + F;
    PYGLM_CUDA(cudaMemsetAsync(ws.colmax, 0, (size_t)(NB + N) * sizeof(unsigned), stream));
    unsigned* cmax = ws.colmax + NB;                                       // [N] largest count per presynaptic column
    const int64_t nbytes = srows * N;
    if (nbytes > 0) {
        const unsigned blocks = (unsigned)std::min<int64_t>(148 * 8, ceil_div(nbytes, 256 * 16));
        tc_spike_colmax_kernel<<<blocks, 256, (size_t)N * sizeof(unsigned), stream>>>(S, nbytes, N, cmax);
        PYGLM_CUDA(cudaGetLastError());
    }
    tc_spike_scales_kernel<<<(unsigned)ceil_div(NS, 128), 128, 0, stream>>>(cmax, d_ibasis, R, B, N, ws.sx);
    PYGLM_CUDA(cudaGetLastError());
    if (F > 0) {
        dim3 gmax((unsigned)std::min<int64_t>(T, 148 * 16), (unsigned)ceil_div(F, 128));
        tc_colmax_kernel<<<gmax, 128, 0, stream>>>(X, T, NS, NB, ldx, ws.colmax);
        PYGLM_CUDA(cudaGetLastError());
        tc_scales_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, stream>>>(ws.colmax, NS, NB, ws.sx);
        PYGLM_CUDA(cudaGetLastError());
    }
    return PYGLM_B200_OK;
}

// Planes from a resident FP32 X (datasets with stimulus features, or planes dropped earlier): scales, then one split pass.
int tc_ensure_planes(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (ws.planes_ready) return PYGLM_B200_OK;
    if (a.X == nullptr) { set_error("tensor-core planes were not built for this dataset"); return PYGLM_B200_ESTATE; }
    const int NB = a.N * a.B + a.F;
    int rc = alloc_planes(ws, a.S, a.T, a.N, a.halo, NB, stream);
    if (rc) return rc;
    if ((rc = tc_compute_scales(ws, a.S, a.T + a.halo, a.N, a.ibasis, a.R, a.B, a.X, a.T, a.F, a.ldx, stream))) return rc;
    tc_split_X_kernel<<<(unsigned)ceil_div((int64_t)a.T * ws.ldp, 256), 256, 0, stream>>>(a.X, a.T, NB, a.ldx, ws.sx, ws.X1, ws.X2, ws.ldp);
    PYGLM_CUDA(cudaGetLastError());
    ws.planes_ready = true;
    return PYGLM_B200_OK;
}

// Single-pass ingest: K1 writes the split planes (and, when `X` is given, the FP32 filtered spike train) straight from its
// gather; no FP32 round trip, no pass over X for the scales.  Used for planes-only datasets and for FP32 datasets without
// stimulus features.
int tc_build_planes_direct(TcWorkspace& ws, const uint8_t* S, int64_t T, int N, int halo, const double* d_ibasis,
                           int R, int B, float* X, int64_t ldx, cudaStream_t stream)
{
    const int NB = N * B;
    int rc = PYGLM_B200_OK;
    if (!ws.X1 && (rc = alloc_planes(ws, S, T, N, halo, NB, stream))) return rc;
    if ((rc = tc_compute_scales(ws, S, T + halo, N, d_ibasis, R, B, nullptr, T, 0, 0, stream))) return rc;
    FilterOut out;
    out.X = X; out.ldx = ldx; out.x_dtype = PYGLM_B200_X_F32;
    out.X1 = ws.X1; out.X2 = ws.X2; out.ldp = ws.ldp; out.sx = ws.sx;
    if ((rc = launch_filter(S, T, N, halo, d_ibasis, R, B, out, stream))) return rc;
    ws.planes_ready = true;
    return PYGLM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// From-spikes evaluation (spikes-only datasets, PYGLM_B200_X_NONE): the operand planes are never resident.  Every
// evaluation walks the recording in time chunks: K1 expands the chunk's spikes straight into the two chunk-sized plane
// buffers (same analytic scales, so the planes are bit-identical to a resident dataset's), the regular tensor-core kernels
// run on the chunk, and the chunk's ll / gradients are added into the outputs in chunk order (deterministic).  A rank
// holding ALL presynaptic spike trains (N*T bytes) can thus evaluate any subset of postsynaptic neurons -- the reference's
// own split (parallel_coord_descent.py:57-157, parallel_gibbs.py:162-168) at sizes whose X would not fit (C4: 164 GB).
// ---------------------------------------------------------------------------------------------
int tc_prepare_streamed(TcWorkspace& ws, const uint8_t* S, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                        cudaStream_t stream)
{
    const int NB = N * B;
    const int64_t ldp = round_up(NB, 8);
    int64_t chunk = std::max<int64_t>(4096, ((int64_t)2048 << 20) / (ldp * 4));                // ~2 GB of planes per chunk
    if (const char* env = getenv("PYGLM_STREAM_CHUNK")) chunk = std::max<int64_t>(1, atoll(env));   // tests: several chunks
    chunk = std::min(round_up(chunk, 128), round_up(T, 128));
    int rc = alloc_planes(ws, S, T, N, halo, NB, stream, chunk);
    if (rc) return rc;
    if ((rc = tc_compute_scales(ws, S, T + halo, N, d_ibasis, R, B, nullptr, T, 0, 0, stream))) return rc;
    ws.streamed = true;
    ws.chunk_rows = chunk;
    ws.planes_ready = true;                       // nothing to build lazily: the planes are produced per evaluation
    return PYGLM_B200_OK;
}

__global__ void tc_accumulate_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

int launch_tc_ll_grad_streamed(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (a.T <= 0 || a.ncols <= 0) return PYGLM_B200_OK;
    if (!ws.streamed || a.F != 0) { set_error("from-spikes evaluation needs a spikes-only dataset without stimulus features"); return PYGLM_B200_ESTATE; }
    const int64_t NB = (int64_t)a.N * a.B;
    const size_t nacc = (size_t)a.ncols * (NB + 2);
    if (ws.acc_elems < nacc) {
        cudaFree(ws.acc);
        ws.acc = nullptr; ws.acc_elems = 0;
        note_allocation();
        PYGLM_CUDA(cudaMalloc(&ws.acc, nacc * sizeof(double)));
        ws.acc_elems = nacc;
    }
    double* t_ll = ws.acc;
    double* t_gb = ws.acc + a.ncols;
    double* t_gw = ws.acc + 2 * (size_t)a.ncols;
    PYGLM_CUDA(cudaMemsetAsync(a.out_ll, 0, a.ncols * sizeof(double), stream));
    if (a.out_gb) PYGLM_CUDA(cudaMemsetAsync(a.out_gb, 0, a.ncols * sizeof(double), stream));
    if (a.out_gw) PYGLM_CUDA(cudaMemsetAsync(a.out_gw, 0, (size_t)a.ncols * NB * sizeof(double), stream));
    uint8_t* const Sp0 = ws.Sp;
    int rc = PYGLM_B200_OK;
    for (int64_t r0 = 0; r0 < a.T && rc == PYGLM_B200_OK; r0 += ws.chunk_rows) {
        const int64_t nt = std::min(ws.chunk_rows, a.T - r0);
        const int64_t h = std::min<int64_t>(a.R, a.halo + r0);                   // left context available for this chunk
        FilterOut fo;
        fo.X1 = ws.X1; fo.X2 = ws.X2; fo.ldp = ws.ldp; fo.sx = ws.sx;
        if ((rc = launch_filter(a.S + (a.halo + r0 - h) * a.N, nt, a.N, (int)h, a.ibasis, a.R, a.B, fo, stream))) break;
        if (ws.map_rows != nt && (rc = tc_encode_maps(ws, nt))) break;
        TcArgs c = a;
        c.X = nullptr; c.T = nt; c.S = a.S + r0 * a.N;
        c.out_ll = t_ll; c.out_gb = a.out_gb ? t_gb : nullptr; c.out_gw = a.out_gw ? t_gw : nullptr;
        ws.Sp = Sp0 + r0 * ws.Np;
        rc = launch_tc_ll_grad(c, ws, stream);
        ws.Sp = Sp0;
        if (rc) break;
        tc_accumulate_kernel<<<(unsigned)ceil_div(a.ncols, 256), 256, 0, stream>>>(a.out_ll, t_ll, a.ncols);
        if (a.out_gb) tc_accumulate_kernel<<<(unsigned)ceil_div(a.ncols, 256), 256, 0, stream>>>(a.out_gb, t_gb, a.ncols);
        if (a.out_gw) tc_accumulate_kernel<<<(unsigned)ceil_div((int64_t)a.ncols * NB, 256), 256, 0, stream>>>(a.out_gw, t_gw, (int64_t)a.ncols * NB);
        PYGLM_CUDA(cudaGetLastError());
    }
    return rc;
}

template <int TAIL, int TILE>
static int launch_fused2(const TcArgs& a, TcWorkspace& ws, TcKernelArgs k, const CUtensorMap* maps, cudaStream_t stream)
{
    using G = V2Geom<TAIL, TILE>;
    k.ntiles = ceil_div(a.T, TILE);
    const int nctas = (int)std::min<int64_t>(k.ntiles, ws.num_sms);
    const int budget = 232448 - 1024;                     // the kernel aligns its carve-up to 1024 bytes itself
    auto slots_for = [&](int rb) { return std::min(kV2MaxSlots, (budget - 1024 - G::kMBytes - rb * G::kRBuf) / G::kSlotBytes); };
    int rbufs = slots_for(1) > slots_for(2) ? 1 : 2;
    if (const char* env = getenv("PYGLM_TC_RBUFS")) rbufs = atoi(env) == 1 ? 1 : 2;
    int nslots = slots_for(rbufs);
    if (const char* env = getenv("PYGLM_TC_SLOTS")) nslots = std::max(4, std::min(nslots, atoi(env)));
    TcV2Args va{k, nslots, rbufs};
    const int smem_bytes = v2_smem_bytes<TAIL, TILE>(nslots, rbufs) + 1024;
    auto kern = a.nlin == PYGLM_B200_NLIN_EXP ? tc_fused2_kernel<PYGLM_B200_NLIN_EXP, TAIL, TILE>
                                              : tc_fused2_kernel<PYGLM_B200_NLIN_SOFTPLUS, TAIL, TILE>;
    PYGLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    // maps: [0,1] 32-feature boxes (64B swizzle), [2,3] 64-feature boxes (128B), [4,5] 16-feature boxes (32B); + 6 for 112-bin boxes
    const int mo = TILE == kTileT ? 0 : 6;
    const CUtensorMap& t1 = TAIL == 16 ? maps[mo + 4] : maps[mo + 0];
    const CUtensorMap& t2 = TAIL == 16 ? maps[mo + 5] : maps[mo + 1];
    kern<<<nctas, kThreads, smem_bytes, stream>>>(maps[mo + 2], maps[mo + 3], t1, t2, va);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

int launch_tc_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream)
{
    if (a.T <= 0 || a.ncols <= 0) return PYGLM_B200_OK;
    const int NB = a.N * a.B + a.F;
    if (!tc_uses_fused_kernel(NB)) return launch_tc_gemm_ll_grad(a, ws, stream);
    int rc = tc_ensure_planes(a, ws, stream);
    if (rc) return rc;
    const int nch = (int)ceil_div(NB, kChunkF);
    const int nmt = (int)ceil_div(nch, 4);
    const int Kp = nch * kChunkF;
    bool v2 = nch >= 4;                                   // 97..160 features: the plane-slot ring kernel
    if (const char* env = getenv("PYGLM_TC_V2")) v2 = v2 && atoi(env) != 0;
    const int tail2 = NB <= 128 ? 0 : (NB <= 144 ? 16 : 32);
    int tile2 = 128;                                      // bins per tile of the V2 kernel (112: a sixth plane slot)
    if (const char* env = getenv("PYGLM_TC_TILE")) tile2 = atoi(env) == 112 ? 112 : 128;
    const int64_t ntiles = ceil_div(a.T, (v2 && tile2 == 112) ? 112 : kTileT);
    const int nctas = (int)std::min<int64_t>(ntiles, ws.num_sms);
    const size_t per_cta = v2 ? (size_t)(128 + tail2) * kNcol + 2 * kNcol : (size_t)nmt * 128 * kNcol + 2 * kNcol;
    if (ws.part_elems < per_cta * nctas) {
        cudaFree(ws.part);
        ws.part = nullptr; ws.part_elems = 0;
        note_allocation();
        PYGLM_CUDA(cudaMalloc(&ws.part, per_cta * nctas * sizeof(double)));
        ws.part_elems = per_cta * nctas;
    }
    const TcSmem L = tc_smem_layout(nch);
    const int smem_bytes = L.total + 1024;
    bool wide = nch >= 4;
    if (const char* env = getenv("PYGLM_TC_WIDE")) wide = wide && atoi(env) != 0;
    auto kern = a.nlin == PYGLM_B200_NLIN_EXP
                    ? (wide ? tc_fused_kernel<PYGLM_B200_NLIN_EXP, true> : tc_fused_kernel<PYGLM_B200_NLIN_EXP, false>)
                    : (wide ? tc_fused_kernel<PYGLM_B200_NLIN_SOFTPLUS, true> : tc_fused_kernel<PYGLM_B200_NLIN_SOFTPLUS, false>);
    PYGLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const CUtensorMap* maps = static_cast<const CUtensorMap*>(ws.tmaps);
    for (int c0 = 0; c0 < a.ncols; c0 += kNcol) {
        const int nc = std::min(kNcol, a.ncols - c0);
        const int n_lo = a.n_lo + c0;
        tc_prep_M_kernel<<<kNcol, 256, 0, stream>>>(a.w, a.A, a.W, a.bias, ws.sx, a.N, a.B, a.F, n_lo, nc, Kp, ws.Mp, ws.colpar);
        PYGLM_CUDA(cudaGetLastError());
        TcKernelArgs k{};
        k.S = a.S; k.T = a.T; k.N = a.N; k.halo = a.halo; k.dt = (float)a.dt; k.nlin = a.nlin;
        k.n_lo = n_lo; k.ncols = nc; k.nch = nch; k.nmt = nmt; k.ntiles = ntiles; k.nfeat = NB;
        k.Sp = ws.Sp; k.Np = ws.Np;
        k.Mp = ws.Mp; k.Kp = Kp; k.colpar = ws.colpar; k.part = ws.part; k.flags = a.flags;
        { const char* dbg = getenv("PYGLM_TC_DEBUG"); k.debug = dbg ? atoi(dbg) : 0; }
        { const char* sl = getenv("PYGLM_TC_SLEEP"); k.producer_sleep_ns = sl ? (unsigned)atoi(sl) : 0u; }
        { const char* fl = getenv("PYGLM_TC_FLUSH"); k.flush = fl ? std::max(1, atoi(fl)) : kFlushTiles; }
        static long long* d_trace = nullptr;
        const bool want_trace = getenv("PYGLM_TC_TRACE") != nullptr;
        constexpr int kTraceWords = 2048 + 16 * 24 * 8;
        if (want_trace && !d_trace) PYGLM_CUDA(cudaMalloc(&d_trace, kTraceWords * sizeof(long long)));
        k.trace = want_trace ? d_trace : nullptr;
        if (v2) {
            int rc2;
            if (tile2 == 112)
                rc2 = tail2 == 0 ? launch_fused2<0, 112>(a, ws, k, maps, stream)
                    : tail2 == 16 ? launch_fused2<16, 112>(a, ws, k, maps, stream) : launch_fused2<32, 112>(a, ws, k, maps, stream);
            else
                rc2 = tail2 == 0 ? launch_fused2<0, 128>(a, ws, k, maps, stream)
                    : tail2 == 16 ? launch_fused2<16, 128>(a, ws, k, maps, stream) : launch_fused2<32, 128>(a, ws, k, maps, stream);
            if (rc2) return rc2;
        } else {
            kern<<<nctas, kThreads, smem_bytes, stream>>>(maps[0], maps[1], maps[2], maps[3], k);
            PYGLM_CUDA(cudaGetLastError());
        }
        if (want_trace) {
            static int dumped = 0;
            if (dumped++ == 5) {
                static long long h[kTraceWords];
                PYGLM_CUDA(cudaStreamSynchronize(stream));
                PYGLM_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
                const long long t0 = h[0];
                if (v2) {       // per-warp phase stamps of tiles 10..12: start, fwd_full, ld done, math done, r_free, r_ready, sums, fold
                    for (int it = 10; it < 13; ++it)
                        for (int w = 0; w < 16; ++w) {
                            const long long* e = h + 2048 + (w * 24 + it) * 8;
                            fprintf(stderr, "tile %d warp %2d: start %7lld | fwd_full +%5lld ld +%4lld math +%5lld gload+r_free +%5lld r_ready +%4lld sums +%4lld fold_add +%5lld\n",
                                    it, w, e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6] - e[5], e[7] - e[6]);
                        }
                }
                fprintf(stderr, "tile  tma_issue | fwd_iss fwd_done_iss bwd_iss bwd_done_iss | epi_start fwd_full r_ready iter_end\n");
                for (int i = 0; i < 24; ++i) {
                    fprintf(stderr, "%3d %9lld | %9lld %9lld %9lld %9lld | %9lld %9lld %9lld %9lld\n", i, h[(0 * 32 + i) * 4] - t0,
                            h[(32 + i) * 4 + 0] - t0, h[(32 + i) * 4 + 1] - t0, h[(32 + i) * 4 + 2] - t0, h[(32 + i) * 4 + 3] - t0,
                            h[(64 + i) * 4 + 0] - t0, h[(64 + i) * 4 + 1] - t0, h[(64 + i) * 4 + 2] - t0, h[(64 + i) * 4 + 3] - t0);
                }
                for (int i = 8; i < 14; ++i)
                    fprintf(stderr, "epi tile %d: fwd_full %lld ld_done +%lld math_done +%lld r_free +%lld r_ready +%lld iter_end +%lld\n", i,
                            h[(64 + i) * 4 + 1] - t0, h[1408 + i * 4 + 0] - h[(64 + i) * 4 + 1], h[1408 + i * 4 + 1] - h[1408 + i * 4 + 0],
                            h[1408 + i * 4 + 2] - h[1408 + i * 4 + 1], h[(64 + i) * 4 + 2] - h[1408 + i * 4 + 2], h[(64 + i) * 4 + 3] - h[(64 + i) * 4 + 2]);
                for (int i = 8; i < 12; ++i)
                    fprintf(stderr, "fold tile %d: start %lld wait +%lld ld +%lld rest +%lld\n", i, h[1536 + i * 4] - t0,
                            h[1536 + i * 4 + 1] - h[1536 + i * 4], h[1536 + i * 4 + 2] - h[1536 + i * 4 + 1], h[1536 + i * 4 + 3] - h[1536 + i * 4 + 2]);
                for (int i = 10; i < 13; ++i) {
                    fprintf(stderr, "r_ready arrivals tile %d:", i);
                    for (int w = kFirstEpiWarp; w < kFirstEpiWarp + kEpiWarps; ++w) fprintf(stderr, " %lld", h[384 + w * 32 + i] - t0);
                    fprintf(stderr, "\n");
                }
            }
        }
        if (v2 && a.ar)
            tc_final2_allreduce_kernel<<<(unsigned)(NB + 1), 32 * 32, 0, stream>>>(
                ws.part, nctas, 128 + tail2, a.N, a.B, a.F, ws.sx, a.A, a.W, *static_cast<const ArEpoch*>(a.ar), a.out_ll);
        else if (v2)
            tc_final2_kernel<<<(unsigned)(NB + 1), 32 * 32, 0, stream>>>(
                ws.part, nctas, 128 + tail2, a.N, a.B, a.F, n_lo, nc, ws.sx, a.A, a.W,
                a.out_ll + c0, a.out_gb ? a.out_gb + c0 : nullptr, a.out_gw ? a.out_gw + (int64_t)c0 * NB : nullptr);
        else
        tc_final_kernel<<<(unsigned)(NB + 1), 32 * kFinalSlices, 0, stream>>>(
            ws.part, nctas, nmt, a.N, a.B, a.F, n_lo, nc, ws.sx, a.A, a.W,
            a.out_ll + c0, a.out_gb ? a.out_gb + c0 : nullptr, a.out_gw ? a.out_gw + (int64_t)c0 * NB : nullptr);
        PYGLM_CUDA(cudaGetLastError());
    }
    return PYGLM_B200_OK;
}

}  // namespace pyglm
