// K1: spike-history filtering.
//   X[t][pre*B+b] = sum_{k=1..R} ibasis[k-1][b] * S[t-k][pre]
// Reference: convolve_with_basis, pyglm/utils/basis.py:201-236 (zero row prepended :220,
// 'full'[:T] :232-234), called from LinearBasisImpulses.preprocess_data (impulse.py:114-130).
//
// The reference runs B dense FFT convolutions.  Spike trains are sparse small integers
// (~2% of bins non-zero), so this kernel gathers instead: a block owns a tile of `tt` output
// bins x `pc` presynaptic columns.
//  * Tile load.  The spike bytes of the tile (plus its R-bin left context) are read with 16-byte
//    vector loads (a tile that spans all N columns is one contiguous byte range) and only the
//    non-zero bytes do anything: they set a bit in a per-column bitmap over the tile's rows
//    (counts > 1 also set a bit in a second bitmap and leave their count in a byte array).
//  * Gather.  A warp task is 32 consecutive output bins of one column: the lanes fetch the bitmap
//    words of the window [first bin - R, last bin), and the warp walks the set bits in increasing
//    time (warp-uniform ffs loop).  Lane l adds basis[lag_l][b] for its own lag; the basis sits in
//    shared memory basis-function-major, so the 32 lanes read 32 consecutive doubles (conflict-free,
//    the 8 bytes per FP64 add that bound this kernel: DESIGN.md section 4).
//    Accumulation is FP64 in increasing spike-time order (the order oracle/convolve_with_basis_direct
//    uses), rounded once to the storage type.
//  * Copy-out.  Results leave through a padded shared-memory tile so global stores are full rows:
//    the full-precision X and / or -- in the same pass -- the FP16 split planes of the tensor-core
//    path (X * sx = X1 + X2 * 2^-11 with the analytic per-feature scales of tc_spike_scales).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace pyglm {

constexpr int kFiltThreads = 512;
constexpr int kFiltWarps = kFiltThreads / 32;
constexpr int kZ = 96;              // leading zeros of each basis function in shared memory (>= 64 trailing ones)

struct FiltLayout {
    int tt, pc, sub, rows, mw, ostride;
    size_t off_basis, off_sx, off_mask, off_cnt, off_out, off_next, total;
};

static FiltLayout filt_layout(int R, int B, int bmax, int rpitch, int tt, int pc, size_t xsz)
{
    FiltLayout L{};
    L.tt = tt; L.pc = pc;
    L.rows = tt + R;
    L.mw = (L.rows + 31) / 32 + 1;
    L.ostride = (pc * B + 2) | 1;                                    // odd: conflict-free staging writes; room for one pad element
    L.sub = 128;
    while (L.sub > 64 && ((size_t)L.sub * L.ostride * xsz > 48 * 1024 || L.sub > tt)) L.sub >>= 1;
    size_t o = 0;
    L.off_basis = o; o += (size_t)rpitch * bmax * sizeof(double);
    L.off_out = o;   o += (size_t)L.sub * L.ostride * xsz;
    o = (o + 15) & ~(size_t)15;
    L.off_mask = o;  o += (size_t)2 * pc * L.mw * sizeof(uint32_t);  // spike bitmap, then the counts>1 bitmap
    L.off_sx = o;    o += (size_t)(pc * B + 2) * sizeof(float);
    L.off_next = o;  o += 8;
    L.off_cnt = o;   o += (size_t)pc * L.rows;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// WX: write the full-precision X; WP: write the split planes.  BEXACT: B == BMAX (no per-basis predicates).
// RP: pitch of one basis function in shared memory (0: R + kZ + 64): kZ zeros, the R values, >= 64 zeros.  A lane whose
// lag falls outside 1..R reads one of the zeros instead of branching (x + 0.0 == x for every x the sums can hold).
template <typename XT, int BMAX, bool BEXACT, bool WX, bool WP, int RP>
__global__ void __launch_bounds__(kFiltThreads)
filter_kernel(const uint8_t* __restrict__ S, int64_t T, int N, int halo,
              const double* __restrict__ ibasis, int R, int B,
              XT* __restrict__ X, int64_t ldx,
              __half* __restrict__ X1, __half* __restrict__ X2, int64_t ldp, const float* __restrict__ sx,
              FiltLayout L)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double*   sB     = reinterpret_cast<double*>(smem + L.off_basis);    // [BMAX][P]   basis-function-major, zero beyond R
    XT*       sOut   = reinterpret_cast<XT*>(smem + L.off_out);          // [sub][ostride]
    uint32_t* sMask  = reinterpret_cast<uint32_t*>(smem + L.off_mask);   // [pc][mw]    bit i: tile row i of the column holds a spike
    uint32_t* sMulti = sMask + L.pc * L.mw;                              // [pc][mw]    ... more than one
    float*    sSx    = reinterpret_cast<float*>(smem + L.off_sx);        // [pc*B]      plane scales of the tile's features
    uint8_t*  sCnt   = smem + L.off_cnt;                                 // [pc][rows]  count, written where sMulti is set
    int*      sNext  = reinterpret_cast<int*>(smem + L.off_next);        // task counter of the gather

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tt = L.tt, pc = L.pc, rows = L.rows, mw = L.mw, ostride = L.ostride;
    const int64_t t0 = (int64_t)blockIdx.x * tt;      // first output bin of the tile
    const int c0 = blockIdx.y * pc;                   // first presynaptic column of the tile
    const int ncol = min(pc, N - c0);
    const int width = ncol * B;

    for (int e = tid; e < 2 * pc * mw; e += kFiltThreads) sMask[e] = 0u;
    const int P = RP ? RP : R + kZ + 64;
    for (int e = tid; e < P * BMAX; e += kFiltThreads) {
        const int b = e / P, k = e - b * P - kZ;
        sB[e] = (b < B && k >= 0 && k < R) ? ibasis[(size_t)k * B + b] : 0.0;
    }
    if (tid == 0) *sNext = 0;
    if (WP) for (int e = tid; e < pc * B + 2; e += kFiltThreads) sSx[e] = e < width ? sx[(size_t)c0 * B + e] : 0.f;
    __syncthreads();

    // ---- tile load: tile row i <-> S row g0 + i; rows outside [0, halo+T) hold no spikes
    const int64_t g0 = (int64_t)halo + t0 - R;
    const int r_lo = (int)max((int64_t)0, -g0);
    const int r_hi = (int)min((int64_t)rows, (int64_t)halo + T - g0);
    auto insert = [&](int row, int col, uint32_t v) {
        atomicOr(&sMask[col * mw + (row >> 5)], 0x80000000u >> (row & 31));      // bit 31 = first row of the word
        if (v > 1u) {
            atomicOr(&sMulti[col * mw + (row >> 5)], 0x80000000u >> (row & 31));
            sCnt[col * rows + row] = (uint8_t)v;
        }
    };
    if (r_hi > r_lo) {
        if (ncol == N) {
            // every column: the tile is one contiguous byte range
            const uint8_t* base = S + (g0 + r_lo) * N;
            const int nbytes = (r_hi - r_lo) * N;
            const int head = min(nbytes, (int)((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15));
            const int nvec = (nbytes - head) >> 4;
            auto insert_flat = [&](int f, uint32_t v) { const int r = f / N; insert(r_lo + r, f - r * N, v); };
            for (int f = tid; f < head; f += kFiltThreads) { const uint32_t v = base[f]; if (v) insert_flat(f, v); }
            const uint4* vp = reinterpret_cast<const uint4*>(base + head);
            for (int v = tid; v < nvec; v += kFiltThreads) {
                const uint4 q = __ldg(vp + v);
                if ((q.x | q.y | q.z | q.w) == 0u) continue;
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (w[i] == 0u) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t b = (w[i] >> (8 * j)) & 0xffu;
                        if (b) insert_flat(head + 16 * v + 4 * i + j, b);
                    }
                }
            }
            for (int f = head + 16 * nvec + tid; f < nbytes; f += kFiltThreads) { const uint32_t v = base[f]; if (v) insert_flat(f, v); }
        } else if (((N | c0 | ncol) & 3) == 0 && (reinterpret_cast<uintptr_t>(S) & 3) == 0) {
            // a block of columns, word-aligned row segments
            const int wpr = ncol >> 2;
            for (int it = tid; it < (r_hi - r_lo) * wpr; it += kFiltThreads) {
                const int r = it / wpr, wi = it - r * wpr;
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(S + (g0 + r_lo + r) * N + c0) + wi);
                if (w == 0u) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t b = (w >> (8 * j)) & 0xffu;
                    if (b) insert(r_lo + r, 4 * wi + j, b);
                }
            }
        } else {
            for (int it = tid; it < (r_hi - r_lo) * ncol; it += kFiltThreads) {
                const int r = it / ncol, c = it - r * ncol;
                const uint32_t b = S[(g0 + r_lo + r) * N + c0 + c];
                if (b) insert(r_lo + r, c, b);
            }
        }
    }
    __syncthreads();

    // copy-out roles, fixed for the whole tile: a thread owns one pair of adjacent features (planes, and X when both are
    // written) or one feature (X alone) and walks down the rows of the sub-tile, `rp` rows per pass of the block
    const int wpl = WP ? ((c0 + ncol == N) ? (int)(ldp - (int64_t)c0 * B) : width) : 0;   // plane columns incl. the zero pad [N*B, ldp)
    const int nitem = WP ? (wpl + 1) / 2 : width;
    const int rp = kFiltThreads / nitem;
    const int prow = tid / nitem, pit = tid - prow * nitem;
    const bool pact = prow < rp;
    const int pe = WP ? 2 * pit : pit;
    const bool pv0 = pe < width, pv1 = WP && pe + 1 < width;
    const float ps0 = (WP && pv0) ? sSx[pe] : 0.f, ps1 = (WP && pv1) ? sSx[pe + 1] : 0.f;   // exact power-of-two scales

    const int sub_n = L.sub, ngrp = sub_n >> 6;
    const uint32_t next_addr = (uint32_t)__cvta_generic_to_shared(sNext);
    for (int sub = 0; sub < tt; sub += sub_n) {
        if (t0 + sub >= T) break;
        // ---- gather: a warp task = 64 consecutive output bins of one column (lane l: bins l and l + 32), handed out
        // by a counter
        for (;;) {
            int task = 0;
            if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(task) : "r"(next_addr) : "memory");
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= ngrp * ncol) break;
            int grp = 0, c = task;
            while (c >= ncol) { c -= ncol; ++grp; }
            const int tb = sub + (grp << 6);            // first output bin of the task within the tile (multiple of 64)
            // lag - 1 of a spike in tile row s is (tb + lane + R - 1) - s for my first bin, 32 more for my second; over the
            // rows of the window's words that is [-94, R + 62]: the basis sits between kZ leading and >= 64 trailing
            // zeros, so no lane needs a range check (x + 0.0 == x)
            const double* lbase = sB + (tb + lane + R - 1 + kZ - 31);
            double acc[BMAX], acc2[BMAX];
#pragma unroll
            for (int b = 0; b < BMAX; ++b) acc[b] = acc2[b] = 0.0;
            const int w0 = tb >> 5, w1 = min((tb + R + 62) >> 5, mw - 1);
            const uint32_t* mrow = sMask + c * mw;
            const uint32_t* xrow = sMulti + c * mw;
            for (int wb = w0; wb <= w1; wb += 32) {
                const int wi = wb + 31 - lane;                          // lane 31 holds the first word: highest ballot bit first
                const uint32_t mm = wi <= w1 ? mrow[wi] : 0u;
                const uint32_t mx = wi <= w1 ? xrow[wi] : 0u;
                uint32_t nz = __ballot_sync(0xffffffffu, mm != 0u);
                const bool anyx = __ballot_sync(0xffffffffu, mx != 0u) != 0u;
                while (nz) {
                    const int j = 31 - __clz(nz);
                    nz ^= 1u << j;
                    uint32_t m = __shfl_sync(0xffffffffu, mm, j);       // bit 31 = first row of the word
                    const double* wbase = lbase - ((wb + 31 - j) << 5);
                    if (!anyx) {                               // the usual case: single spikes, no multiply
                        do {
                            const int hb = 31 - __clz(m);      // warp-uniform: the spike sits in row 31 - hb of the word
                            m ^= 1u << hb;
                            const double* p = wbase + hb;
#pragma unroll
                            for (int b = 0; b < BMAX; ++b)
                                if (BEXACT || b < B) {
                                    acc[b] = __dadd_rn(acc[b], p[b * P]);
                                    acc2[b] = __dadd_rn(acc2[b], p[b * P + 32]);
                                }
                        } while (m);
                    } else {                                   // rare: some bin of these words holds several spikes
                        const uint32_t x = __shfl_sync(0xffffffffu, mx, j);
                        do {
                            const int hb = 31 - __clz(m);
                            m ^= 1u << hb;
                            const double* p = wbase + hb;
                            const double dc = (x >> hb) & 1u ? (double)sCnt[c * rows + ((wb + 31 - j) << 5) + 31 - hb] : 1.0;
#pragma unroll
                            for (int b = 0; b < BMAX; ++b)
                                if (BEXACT || b < B) {
                                    acc[b] = __dadd_rn(acc[b], __dmul_rn(dc, p[b * P]));
                                    acc2[b] = __dadd_rn(acc2[b], __dmul_rn(dc, p[b * P + 32]));
                                }
                        } while (m);
                    }
                }
            }
            XT* o = sOut + ((grp << 6) + lane) * ostride + c * B;
#pragma unroll
            for (int b = 0; b < BMAX; ++b)
                if (BEXACT || b < B) { o[b] = (XT)acc[b]; o[32 * ostride + b] = (XT)acc2[b]; }
        }
        __syncthreads();
        if (tid == 0) *sNext = 0;
        // ---- coalesced copy-out of the valid part of the sub-tile
        const int nrow = (int)min((int64_t)sub_n, T - (t0 + sub));
        if (pact) {
            const XT* src = sOut + prow * ostride + pe;
            const int64_t t = t0 + sub + prow;
            if (WP) {
                __half2* d1 = reinterpret_cast<__half2*>(X1 + t * ldp + (int64_t)c0 * B) + pit;
                __half2* d2 = reinterpret_cast<__half2*>(X2 + t * ldp + (int64_t)c0 * B) + pit;
                XT* dx = WX ? X + t * ldx + (int64_t)c0 * B + pe : nullptr;
                for (int r = prow; r < nrow; r += rp) {
                    const float x0 = pv0 ? (float)src[0] : 0.f, x1 = pv1 ? (float)src[1] : 0.f;
                    if (WX) {
                        if (pv1) *reinterpret_cast<float2*>(dx) = make_float2(x0, x1);
                        else if (pv0) *reinterpret_cast<float*>(dx) = x0;
                        dx += (int64_t)rp * ldx;
                    }
                    const float v0 = x0 * ps0, v1 = x1 * ps1;
                    const __half2 h1 = __floats2half2_rn(v0, v1);
                    const float2 f1 = __half22float2(h1);
                    const __half2 h2 = __floats2half2_rn((v0 - f1.x) * 2048.0f, (v1 - f1.y) * 2048.0f);
                    __stcs(reinterpret_cast<unsigned*>(d1), *reinterpret_cast<const unsigned*>(&h1));     // written once, read by a
                    __stcs(reinterpret_cast<unsigned*>(d2), *reinterpret_cast<const unsigned*>(&h2));     // later kernel: streaming
                    d1 += (int64_t)rp * (ldp >> 1);
                    d2 += (int64_t)rp * (ldp >> 1);
                    src += rp * ostride;
                }
            } else {
                XT* dx = X + t * ldx + (int64_t)c0 * B + pe;
                for (int r = prow; r < nrow; r += rp) {
                    *dx = *src;
                    dx += (int64_t)rp * ldx;
                    src += rp * ostride;
                }
            }
        }
        __syncthreads();
    }
}

template <typename XT, int BMAX, bool BEXACT, bool WX, bool WP, int RP>
static int launch_filter_r(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                           const FilterOut& out, cudaStream_t stream)
{
    // the largest tile with two resident blocks per SM, else the largest that fits at all
    const int cand[][2] = {{512, 32}, {256, 32}, {128, 32}, {64, 32}, {64, 16}, {64, 8}, {64, 4}, {64, 2}, {64, 1}};
    FiltLayout L{};
    bool found = false;
    int tt_max = 1 << 30;
    if (const char* env = getenv("PYGLM_FILT_TT")) tt_max = std::max(64, atoi(env));       // experiments: cap the tile length
    for (size_t limit : {(size_t)113 * 1024, (size_t)227 * 1024}) {
        for (auto& c : cand) {
            if (c[0] > tt_max) continue;
            L = filt_layout(R, B, BMAX, RP ? RP : R + kZ + 64, c[0], std::min(c[1], N), sizeof(XT));
            if (L.total <= limit) { found = true; break; }
        }
        if (found) break;
    }
    if (!found) {
        set_error("filter: R=%d B=%d needs more shared memory than one SM has", R, B);
        return PYGLM_B200_EUNSUPPORTED;
    }
    auto kern = filter_kernel<XT, BMAX, BEXACT, WX, WP, RP>;
    PYGLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    dim3 grid((unsigned)ceil_div(T, L.tt), (unsigned)ceil_div(N, L.pc));
    kern<<<grid, kFiltThreads, L.total, stream>>>(dS, T, N, halo, d_ibasis, R, B, static_cast<XT*>(out.X), out.ldx,
                                                   out.X1, out.X2, out.ldp, out.sx, L);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

template <typename XT, int BMAX, bool BEXACT, bool WX, bool WP>
static int launch_filter_v(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                           const FilterOut& out, cudaStream_t stream)
{
    if (R + kZ + 64 <= 360) return launch_filter_r<XT, BMAX, BEXACT, WX, WP, 360>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    return launch_filter_r<XT, BMAX, BEXACT, WX, WP, 0>(dS, T, N, halo, d_ibasis, R, B, out, stream);
}

template <typename XT, bool WX, bool WP>
static int launch_filter_b(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                           const FilterOut& out, cudaStream_t stream)
{
    if (B == 5)  return launch_filter_v<XT, 5, true, WX, WP>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    if (B == 10) return launch_filter_v<XT, 10, true, WX, WP>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    if (B <= 4)  return launch_filter_v<XT, 4, false, WX, WP>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    if (B <= 8)  return launch_filter_v<XT, 8, false, WX, WP>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    return launch_filter_v<XT, 16, false, WX, WP>(dS, T, N, halo, d_ibasis, R, B, out, stream);
}

int launch_filter(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                  const FilterOut& out, cudaStream_t stream)
{
    if (B < 1 || B > kMaxBasis) {
        set_error("filter: B=%d outside [1,%d]", B, kMaxBasis);
        return PYGLM_B200_EUNSUPPORTED;
    }
    if (T <= 0) return PYGLM_B200_OK;
    const bool wx = out.X != nullptr, wp = out.X1 != nullptr;
    if (!wx && !wp) return PYGLM_B200_OK;
    if (wp && (out.X2 == nullptr || out.sx == nullptr || (out.ldp & 1) || out.ldp < (int64_t)N * B)) {
        set_error("filter: bad plane outputs");
        return PYGLM_B200_EINVAL;
    }
    if (wx && out.x_dtype == PYGLM_B200_X_F64) {
        if (wp) { set_error("filter: split planes are built from the FP32 filtered spike train"); return PYGLM_B200_EUNSUPPORTED; }
        return launch_filter_b<double, true, false>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    }
    if (wx && wp) return launch_filter_b<float, true, true>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    if (wx)       return launch_filter_b<float, true, false>(dS, T, N, halo, d_ibasis, R, B, out, stream);
    return launch_filter_b<float, false, true>(dS, T, N, halo, d_ibasis, R, B, out, stream);
}

// ---------------------------------------------------------------------------------
// St[n][t] = S[halo+t][n]: column-major copy of the spikes for the per-column streams
// the Gibbs kernel reads (glm.py:52 indexes S[:, n], a strided column).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_spikes_kernel(const uint8_t* __restrict__ S, int64_t T, int N, int halo, uint8_t* __restrict__ St)
{
    __shared__ uint8_t tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t t = t0 + r;
        const int n = n0 + tx;
        tile[r][tx] = (t < T && n < N) ? S[((int64_t)halo + t) * N + n] : (uint8_t)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r;
        const int64_t t = t0 + tx;
        if (n < N && t < T) St[(int64_t)n * T + t] = tile[tx][r];
    }
}

int launch_transpose_spikes(const uint8_t* dS, int64_t T, int N, int halo, uint8_t* dSt, cudaStream_t stream)
{
    if (T <= 0) return PYGLM_B200_OK;
    dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(N, 32));
    transpose_spikes_kernel<<<grid, 256, 0, stream>>>(dS, T, N, halo, dSt);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

// ---------------------------------------------------------------------------------
// Stimulus features (BasisStimulus, pyglm/components/bkgd.py:122-154): the interpolated stimulus is real
// valued, so its basis projection is a dense causal convolution in FP64,
//   out[t][d*B+b] = sum_{k=1..R} ibasis[k-1][b] * stim[t-k][d]          (utils/basis.py:212-236)
// summed in increasing lag.  One block = 256 bins of one stimulus dimension; window and basis in smem.
// ---------------------------------------------------------------------------------
constexpr int kDenseBins = 256;

__global__ void __launch_bounds__(kDenseBins)
filter_dense_kernel(const double* __restrict__ stim, int64_t T, int D, const double* __restrict__ ibasis, int R, int B,
                    double* __restrict__ out)
{
    extern __shared__ double dsm[];
    double* sIb = dsm;                       // [R][B]
    double* sS = dsm + (size_t)R * B;        // [R + kDenseBins]: stim[t0-R .. t0+255][d]
    const int d = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * kDenseBins;
    for (int i = threadIdx.x; i < R * B; i += kDenseBins) sIb[i] = ibasis[i];
    for (int i = threadIdx.x; i < R + kDenseBins; i += kDenseBins) {
        const int64_t t = t0 - R + i;
        sS[i] = (t >= 0 && t < T) ? stim[t * D + d] : 0.0;
    }
    __syncthreads();
    const int64_t t = t0 + threadIdx.x;
    if (t >= T) return;
    double acc[kMaxBasis];
#pragma unroll
    for (int b = 0; b < kMaxBasis; ++b) acc[b] = 0.0;
    for (int k = 1; k <= R; ++k) {
        const double v = sS[R + threadIdx.x - k];
#pragma unroll
        for (int b = 0; b < kMaxBasis; ++b)
            if (b < B) acc[b] = fma(sIb[(k - 1) * B + b], v, acc[b]);
    }
#pragma unroll
    for (int b = 0; b < kMaxBasis; ++b)
        if (b < B) out[t * ((int64_t)D * B) + (int64_t)d * B + b] = acc[b];
}

int launch_filter_dense(const double* d_stim, int64_t T, int D, const double* d_ibasis, int R, int B, double* d_out,
                        cudaStream_t stream)
{
    if (T <= 0 || D <= 0) return PYGLM_B200_OK;
    const size_t smem = ((size_t)R * B + R + kDenseBins) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("filter_dense: R=%d x B=%d basis does not fit in shared memory", R, B);
        return PYGLM_B200_EUNSUPPORTED;
    }
    PYGLM_CUDA(cudaFuncSetAttribute(filter_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, kDenseBins), (unsigned)D);
    filter_dense_kernel<<<grid, kDenseBins, smem, stream>>>(d_stim, T, D, d_ibasis, R, B, d_out);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

// X[t][col0 + f] = fstim[t][f]: the stimulus features ride behind the spike-history features of X
template <typename XT>
__global__ void __launch_bounds__(256)
fill_stim_kernel(const double* __restrict__ fstim, int64_t T, int F, XT* __restrict__ X, int64_t ldx, int64_t col0)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * F) return;
    const int64_t t = idx / F;
    const int f = (int)(idx - t * F);
    X[t * ldx + col0 + f] = (XT)fstim[idx];
}

int launch_fill_stim(const double* d_fstim, int64_t T, int F, void* dX, int64_t ldx, int64_t col0, int x_dtype,
                     cudaStream_t stream)
{
    if (T <= 0 || F <= 0) return PYGLM_B200_OK;
    const unsigned blocks = (unsigned)ceil_div(T * F, 256);
    if (x_dtype == PYGLM_B200_X_F32) fill_stim_kernel<float><<<blocks, 256, 0, stream>>>(d_fstim, T, F, (float*)dX, ldx, col0);
    else fill_stim_kernel<double><<<blocks, 256, 0, stream>>>(d_fstim, T, F, (double*)dX, ldx, col0);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

}  // namespace pyglm
