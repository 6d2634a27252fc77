// K4: batched delta-log-likelihood for the collapsed spike-and-slab Gibbs sampler over A/W.
//
// Reference: CollapsedGibbsNetworkColumnUpdate (pyglm/inference/gibbs.py:775-1250).  For one
// edge (n_pre -> n) it evaluates `_glm_ll` (:910-937) at 10 Gauss-Hermite abscissae plus w=0
// (:1002-1032), each a full pass over T bins through Theano, and rebuilds I_other with a
// T x N gemv before every edge (:835-864).
//
// Here one launch evaluates every candidate weight of every edge in the batch:
//   u[t]      = sum_b X[t][pre*B+b] * w[col][pre*B+b]          (I_imp[:, n_pre], impulse.py:58)
//   I_other   = I_net[col][t] - (A W)[pre][col] * u[t]         (gibbs.py:838-861)
//   x_q[t]    = bias[col] + I_other + w_q * u[t]               (gibbs.py:914, glm.py:43-45)
//   ll[m][q]  = sum_t -dt f(x_q) + S[t][col] log f(x_q)        (glm.py:52)
// base current, u and the spike are read once per bin and all Q candidates stay in registers.
// Everything is FP64 with a fixed reduction order: the sampler's accept/reject decision
// compares differences of these sums against logit(u), so results must be reproducible.
#include <math_constants.h>

#include "common.cuh"

namespace pyglm {

#ifndef PYGLM_GIBBS_MINBLOCKS
#define PYGLM_GIBBS_MINBLOCKS 2      // 128 registers: two resident blocks per SM
#endif
constexpr int kGibbsThreads = 256;
constexpr int kGibbsChunk = 8192;     // bins per block

int gibbs_num_chunks(int64_t T) { return (int)ceil_div(T, kGibbsChunk); }

// The per-bin rate of the softplus model is lam = max(x,0) + L(|x|), L(a) = log1p(e^-a).  The sampler evaluates
// it (Q+1) T times per edge.  The library exp + log1p pair costs ~95 instructions per evaluation and its
// branches keep the Q+1 candidates of a bin from overlapping, so L is computed by range instead
// (absolute error < 1e-16 everywhere):
//   a >= 5.5 : e = e^-a by Cody-Waite reduction and a degree-9 polynomial (coefficients in the constant bank),
//              then L = e q, q = 1 - e/2 + e^2/3 - ... - e^5/6  (e < 4.1e-3).  ~21 FP64 operations, no branch and
//              no memory access: the candidates' dependency chains interleave and keep the FP64 pipe busy.
//   a < 5.5  : 23 intervals of width 1/4 centred on a = i/4, a degree-9 polynomial in z = 8(a - i/4) on each,
//              coefficients in shared memory.  Applied as a fix-up after the branch-free pass, only in warps
//              where some lane needs it.
// Both polynomial sets are Chebyshev interpolants expanded in monomials, built once on the host in long double.
constexpr int kSpDeg = 9;
constexpr int kSpRows = 23;
constexpr int kSpExpDeg = 9;
constexpr double kSpSplit = 5.5;
constexpr double kSpMagic = 6755399441055744.0;          // 1.5 * 2^52: adding it rounds to the nearest integer
__device__ double g_softplus_tab[kSpRows * (kSpDeg + 1)];
__constant__ double c_sp_exp[kSpExpDeg + 1];

// e = e^-|x| and q with L(|x|) = e q, for |x| in [5.5, inf) (finite garbage below 5.5; callers replace it).
// FP64 min/max/compare are multi-instruction on this part, so the range handling stays on the integer pipe:
// the exponent n is clamped (e^-|x| flushes towards zero past |x| = 708) and 2^n is added into the exponent field.
__device__ __forceinline__ void sp_exp_series(double x, double& e, double& q)
{
    const double t = fma(fabs(x), -1.4426950408889634, kSpMagic);
    const int n = max(__double2loint(t), -1022);          // rint(-|x| log2(e))
    const double fn = t - kSpMagic;
    double r = fma(fn, -6.93147180369123816490e-01, -fabs(x));
    r = fma(fn, -1.90821492927058770002e-10, r);          // r = -|x| - n ln2, |r| <= ln2/2
    double p = c_sp_exp[kSpExpDeg];
#pragma unroll
    for (int k = kSpExpDeg - 1; k >= 0; --k) p = fma(p, r, c_sp_exp[k]);
    e = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));       // p in [0.70, 1.42): p 2^n
    q = fma(e, -1.0 / 6.0, 0.2);
    q = fma(q, e, -0.25);
    q = fma(q, e, 1.0 / 3.0);
    q = fma(q, e, -0.5);
    q = fma(q, e, 1.0);
}

// max(x, 0) and |x| < 5.5 from the high word (integer pipe)
__device__ __forceinline__ double sp_relu(double x)
{
    const int hi = __double2hiint(x);
    return __hiloint2double(hi < 0 ? 0 : hi, hi < 0 ? 0 : __double2loint(x));
}
__device__ __forceinline__ bool sp_small(double x) { return (__double2hiint(x) & 0x7fffffff) < 0x40160000; }

// L(a) for a < 5.5 from the shared-memory table
__device__ __forceinline__ double sp_table(double a, const double* __restrict__ tab)
{
    const double tm = fma(a, 4.0, kSpMagic);
    const int i = __double2loint(tm);                     // rint(4a) in 0..22
    const double fi = tm - kSpMagic;
    const double z = fma(a, 8.0, -2.0 * fi);              // in [-1, 1]
    const double* __restrict__ r = tab + i * (kSpDeg + 1);
    double p = r[kSpDeg];
#pragma unroll
    for (int k = kSpDeg - 1; k >= 0; --k) p = fma(p, z, r[k]);
    return p;
}

// log(softplus(x)) for one value (spike bins).  For x <= -5.5, lam = e^x q exactly, so log(lam) = x + log(q)
// with log(q) from the series of log1p(q - 1): full relative accuracy without an FP64 log.  Where e^x
// underflows the reference formula gives log(0) = -inf (nlin.py:43, glm.py:52), and so does this.
__device__ __forceinline__ double log_softplus(double x, const double* __restrict__ tab)
{
    if (sp_small(x)) return log(sp_relu(x) + sp_table(fabs(x), tab));
    double e, q;
    sp_exp_series(x, e, q);
    if (x > 0.0) return log(fma(e, q, x));
    if (x < -745.1332191019411) return -CUDART_INF;
    const double d = q - 1.0;                             // |d| < 2.1e-3
    double l = fma(d, 0.2, -0.25);
    l = fma(l, d, 1.0 / 3.0);
    l = fma(l, d, -0.5);
    l = fma(l, d, 1.0);
    return fma(l, d, x);
}

// Chebyshev interpolant of f on [c-h, c+h] of degree n-1, returned as monomial coefficients in (x - c)/h
template <int n, typename Fn>
static void cheb_monomial(Fn f, long double c, long double h, long double* out)
{
    long double Tm[n][n] = {};                            // Chebyshev polynomials as monomial coefficient rows
    Tm[0][0] = 1;
    if (n > 1) Tm[1][1] = 1;
    for (int k = 2; k < n; ++k)
        for (int j = 0; j <= k; ++j) Tm[k][j] = (j ? 2 * Tm[k - 1][j - 1] : 0) - Tm[k - 2][j];
    const long double pi = acosl(-1.0L);
    long double fv[n], cheb[n];
    for (int k = 0; k < n; ++k) fv[k] = f(c + h * cosl(pi * (k + 0.5L) / n));
    for (int j = 0; j < n; ++j) {
        long double acc = 0;
        for (int k = 0; k < n; ++k) acc += fv[k] * cosl(pi * j * (k + 0.5L) / n);
        cheb[j] = acc * (j ? 2.0L : 1.0L) / n;
    }
    for (int j = 0; j < n; ++j) {
        long double m = 0;
        for (int k = j; k < n; ++k) m += cheb[k] * Tm[k][j];
        out[j] = m;
    }
}

static int ensure_softplus_table()
{
    static bool done[64] = {};
    int dev = 0;
    PYGLM_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return PYGLM_B200_OK;
    constexpr int n = kSpDeg + 1;
    static double tab[kSpRows * n];
    long double m[16];
    for (int i = 0; i < kSpRows; ++i) {
        cheb_monomial<n>([](long double a) { return log1pl(expl(-a)); }, 0.25L * i, 0.125L, m);
        for (int j = 0; j < n; ++j) tab[i * n + j] = (double)m[j];
    }
    double ce[kSpExpDeg + 1];
    const long double h = 0.3470L;                        // |r| <= ln2/2 = 0.34657...
    cheb_monomial<kSpExpDeg + 1>([](long double r) { return expl(r); }, 0.0L, h, m);
    long double hp = 1;
    for (int j = 0; j <= kSpExpDeg; ++j) { ce[j] = (double)(m[j] / hp); hp *= h; }
    PYGLM_CUDA(cudaMemcpyToSymbol(g_softplus_tab, tab, sizeof(tab)));
    PYGLM_CUDA(cudaMemcpyToSymbol(c_sp_exp, ce, sizeof(ce)));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return PYGLM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// From-spikes mode: u[t] = I_imp[t, pre] = sum_b X[t, pre, b] w_b = sum_{k=1..R} h[k-1] S[t-k][pre],  h = ibasis . w
// (utils/basis.py:201-236 folded into impulse.py:58).  X (T x N x B) is never read: one byte per bin of the
// presynaptic spike train instead of B values, which is what lets a GPU hold a population whose filtered spike train
// would not fit (C4: 164 GB of X, 4 GB of spikes).  Spikes are sparse, so the block compacts the spikes of its chunk
// plus the R-bin left context into a time-ordered list and records, for every bin of the window, how many spikes
// precede it; a bin then sums c * h[lag] over the ~4 list entries of its own window, in increasing spike time (a fixed
// order: reproducible sums), with no barrier and no atomics.
// Shared memory: sh [R] doubles | spos [W] u16 | scum [W + 1] u16 | scnt [W] u8,  W = kGibbsChunk + R.
// ---------------------------------------------------------------------------------------------
struct SpkSmem {
    double* sh; unsigned short* spos; unsigned short* scum; unsigned char* scnt;
    int R;
};
static inline size_t spk_smem_bytes(int R)
{
    const size_t W = (size_t)kGibbsChunk + R;
    return (size_t)round_up(R, 2) * 8 + round_up(W * 2, 16) + round_up((W + 1) * 2, 16) + round_up(W, 16);
}
__device__ inline SpkSmem spk_carve(unsigned char* base, int R)
{
    SpkSmem m;
    const size_t W = (size_t)kGibbsChunk + R;
    m.sh = reinterpret_cast<double*>(base);
    m.spos = reinterpret_cast<unsigned short*>(m.sh + ((R + 1) & ~1));
    m.scum = m.spos + ((W * 2 + 15) & ~(size_t)15) / 2;
    m.scnt = reinterpret_cast<unsigned char*>(m.scum) + (((W + 1) * 2 + 15) & ~(size_t)15);
    m.R = R;
    return m;
}

// Prepare the gather for the bins [tbeg, tbeg + nbins) of presynaptic row `st_pre` (indexable from -halo):
// h = scale * ibasis . w, the spike list of the window [tbeg - R, tbeg + nbins - 1) and the per-position spike counts.
// All threads of the block must call it; it ends with a barrier.
__device__ void spk_build(const SpkSmem& m, const uint8_t* __restrict__ st_pre, int halo, int64_t tbeg, int nbins,
                          const double* __restrict__ ibasis, int B, const double* __restrict__ w, double scale,
                          int* s_scan /* [kGibbsThreads / 32 + 1] */)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = m.R;
    __syncthreads();                                      // the previous edge's gathers are done with these arrays
    for (int l = tid; l < R; l += kGibbsThreads) {
        double h = 0.0;
        for (int b = 0; b < B; ++b) h = fma(ibasis[(int64_t)l * B + b], w[b], h);
        m.sh[l] = h * scale;
    }
    // each thread scans a contiguous segment of the window, so the list keeps time order
    const int W = R + nbins - 1;
    const int seg = (W + kGibbsThreads - 1) / kGibbsThreads;
    const int p0 = min(W, tid * seg), p1 = min(W, p0 + seg);
    int cnt = 0;
    for (int p = p0; p < p1; ++p) {
        const int64_t t = tbeg - R + p;
        cnt += (t >= -(int64_t)halo && st_pre[t] != 0) ? 1 : 0;
    }
    int incl = cnt;                                       // block-wide exclusive scan of the counts
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int i = 0; i < kGibbsThreads / 32; ++i) { const int v = s_scan[i]; s_scan[i] = run; run += v; }
        s_scan[kGibbsThreads / 32] = run;
    }
    __syncthreads();
    int k = s_scan[warp] + incl - cnt;
    for (int p = p0; p < p1; ++p) {
        const int64_t t = tbeg - R + p;
        const unsigned char c = t >= -(int64_t)halo ? st_pre[t] : (unsigned char)0;
        m.scum[p] = (unsigned short)k;                    // spikes at window positions < p
        if (c) { m.spos[k] = (unsigned short)p; m.scnt[k] = c; ++k; }
    }
    if (tid == kGibbsThreads - 1) m.scum[W] = (unsigned short)s_scan[kGibbsThreads / 32];
    __syncthreads();
}

// u of bin tbeg + i: the spikes at window positions [i, i + R) are lags R .. 1 away (position p is bin tbeg - R + p)
__device__ __forceinline__ double spk_u(const SpkSmem& m, int i)
{
    const int k0 = m.scum[i], k1 = m.scum[i + m.R];
    double u = 0.0;
    for (int k = k0; k < k1; ++k) u = fma((double)m.scnt[k], m.sh[i + m.R - 1 - (int)m.spos[k]], u);
    return u;
}

// log(lam) of the Poisson term is needed only in the ~2% of bins that hold a spike.  Those bins are handed to
// the whole warp: lane q evaluates candidate q, so one FP64 log serves all candidates.
template <typename XT, int QMAX, int NLIN, int BMAX, bool SPK>
__global__ void __launch_bounds__(kGibbsThreads, PYGLM_GIBBS_MINBLOCKS)
gibbs_delta_kernel(GibbsArgs g, const int32_t* __restrict__ cols, const int32_t* __restrict__ pres,
                   int Q, const double* __restrict__ wcand)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ int sScan[kGibbsThreads / 32 + 1];
    __shared__ double sW[kMaxBasis];
    __shared__ double sRed[QMAX][kGibbsThreads / 32];
    __shared__ double sRedSp[kGibbsThreads / 32][32];
    __shared__ double sTab[NLIN == PYGLM_B200_NLIN_SOFTPLUS ? kSpRows * (kSpDeg + 1) : 1];

    const XT* __restrict__ X = static_cast<const XT*>(g.X);
    const int m = blockIdx.y;
    const int col = cols[m], pre = pres[m];
    const int nl = col - g.n_lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t NB = (int64_t)g.N * g.B + g.F;           // row pitch of w: all features

    if (tid < kMaxBasis) sW[tid] = tid < g.B ? g.w[(int64_t)col * NB + (int64_t)pre * g.B + tid] : 0.0;
    if (NLIN == PYGLM_B200_NLIN_SOFTPLUS)
        for (int i = tid; i < kSpRows * (kSpDeg + 1); i += kGibbsThreads) sTab[i] = g_softplus_tab[i];
    __syncthreads();

    const double aw_old = (double)g.A[(int64_t)pre * g.N + col] * g.W[(int64_t)pre * g.N + col];
    const double bias = g.bias[col];
    double wq[QMAX], acc[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        wq[q] = (q < Q) ? wcand[(int64_t)m * Q + q] : 0.0;
        acc[q] = 0.0;
    }
    const double w_lane = (lane < Q) ? wcand[(int64_t)m * Q + lane] : 0.0;   // candidate `lane` (spike path)
    double acc_sp = 0.0;

    const int64_t tbeg = (int64_t)blockIdx.x * kGibbsChunk;
    const int64_t tend = min(g.T, tbeg + kGibbsChunk);
    const double* __restrict__ inet = g.Inet + (int64_t)nl * g.T;
    const uint8_t* __restrict__ st = g.St + (int64_t)col * g.ldst;
    const XT* __restrict__ xcol = SPK ? nullptr : X + (int64_t)pre * g.B * g.T;      // feature-major copy: Xt[j][t]
    SpkSmem spk{};
    if (SPK) {                                             // spike list of the presynaptic train for this chunk
        spk = spk_carve(dyn_smem, g.R);
        spk_build(spk, g.St + (int64_t)pre * g.ldst, g.halo, tbeg, (int)(tend - tbeg), g.ibasis, g.B, sW, 1.0, sScan);
    }

    // the operands of the next 256 bins are fetched while the current ones are evaluated
    XT xr[BMAX];
    double ir;
    unsigned sr;
#define PYGLM_GIBBS_FETCH(TT)                                                                   \
    do {                                                                                        \
        const int64_t tt_ = (TT);                                                               \
        const bool lv_ = tt_ < tend;                                                            \
        if (!SPK) {                                                                             \
            _Pragma("unroll") for (int b = 0; b < BMAX; ++b)                                    \
                xr[b] = (lv_ && b < g.B) ? xcol[(int64_t)b * g.T + tt_] : (XT)0;                \
        }                                                                                       \
        ir = lv_ ? inet[tt_] : 0.0;                                                             \
        sr = lv_ ? (unsigned)st[tt_] : 0u;                                                      \
    } while (0)
    PYGLM_GIBBS_FETCH(tbeg + tid);

    for (int64_t t0 = tbeg; t0 < tend; t0 += kGibbsThreads) {        // warp-uniform trip count
        const bool live = t0 + tid < tend;
        double u = 0.0;
        if (SPK) {
            u = live ? spk_u(spk, (int)(t0 - tbeg) + tid) : 0.0;
        } else {
#pragma unroll
            for (int b = 0; b < BMAX; ++b) u = fma((double)xr[b], sW[b], u);
        }
        const double base = bias + (ir - aw_old * u);
        const double s = (double)sr;
        PYGLM_GIBBS_FETCH(t0 + kGibbsThreads + tid);

        if (NLIN == PYGLM_B200_NLIN_SOFTPLUS) {
            bool small = false;
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {                         // branch-free pass: Q independent chains
                const double x = fma(wq[q], u, base);
                double e, qq;
                sp_exp_series(x, e, qq);
                const bool sm = sp_small(x);
                small |= sm;
                acc[q] = fma((live && !sm) ? 1.0 : 0.0, fma(e, qq, sp_relu(x)), acc[q]);
            }
            if (__any_sync(0xffffffffu, small)) {                    // second pass: |x| < 5.5 from the table
#pragma unroll
                for (int q = 0; q < QMAX; ++q) {
                    const double x = fma(wq[q], u, base);
                    if (live && sp_small(x)) acc[q] += sp_relu(x) + sp_table(fabs(x), sTab);
                }
            }
            // spike bins (~2%): one lane per candidate evaluates log(lam) for the whole warp
            unsigned mask = __ballot_sync(0xffffffffu, live && s != 0.0);   // (sr already holds the next bin)
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double bs = __shfl_sync(0xffffffffu, base, src);
                const double us = __shfl_sync(0xffffffffu, u, src);
                const double ss = __shfl_sync(0xffffffffu, s, src);
                if (lane < Q) acc_sp += ss * log_softplus(fma(w_lane, us, bs), sTab);
            }
        } else {
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                const double x = fma(wq[q], u, base);
                acc[q] = fma(live ? 1.0 : 0.0, -g.dt * exp(x) + x * s, acc[q]);  // exp nonlinearity: log(lam) = x
            }
        }
    }
#undef PYGLM_GIBBS_FETCH
    // block reduction, fixed order: lanes (xor tree), then warps 0..7; the spike sums live in lane q
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        double v = NLIN == PYGLM_B200_NLIN_SOFTPLUS ? -g.dt * acc[q] : acc[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sRed[q][warp] = v;
    }
    sRedSp[warp][lane] = acc_sp;
    __syncthreads();
    if (tid < Q) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kGibbsThreads / 32; ++w) v += sRed[tid][w] + sRedSp[w][tid];
        g.partial[((int64_t)m * g.nchunks + blockIdx.x) * Q + tid] = v;
    }
}

__global__ void gibbs_reduce_kernel(const double* __restrict__ partial, int M, int nchunks, int Q,
                                    double* __restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * Q) return;
    const int m = idx / Q, q = idx - m * Q;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[((int64_t)m * nchunks + c) * Q + q];
    out[idx] = s;
}

template <typename XT, bool SPK>
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_commit_kernel(GibbsArgs g, const int32_t* __restrict__ cols, const int32_t* __restrict__ pres,
                    const int8_t* __restrict__ anew, const double* __restrict__ wnew)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ int sScan[kGibbsThreads / 32 + 1];
    __shared__ double sW[kMaxBasis];
    const XT* __restrict__ X = static_cast<const XT*>(g.X);
    const int m = blockIdx.y;
    const int col = cols[m], pre = pres[m];
    const int nl = col - g.n_lo;
    const int tid = threadIdx.x;
    const int64_t NB = (int64_t)g.N * g.B + g.F;           // row pitch of w: all features
    if (tid < g.B) sW[tid] = g.w[(int64_t)col * NB + (int64_t)pre * g.B + tid];
    __syncthreads();
    const double aw_old = (double)g.A[(int64_t)pre * g.N + col] * g.W[(int64_t)pre * g.N + col];
    const double aw_new = (double)anew[m] * wnew[m];
    const double delta = aw_new - aw_old;
    if (delta == 0.0) return;
    const int64_t tbeg = (int64_t)blockIdx.x * kGibbsChunk;
    const int64_t tend = min(g.T, tbeg + kGibbsChunk);
    double* __restrict__ inet = g.Inet + (int64_t)nl * g.T;
    if (SPK) {
        SpkSmem spk = spk_carve(dyn_smem, g.R);
        spk_build(spk, g.St + (int64_t)pre * g.ldst, g.halo, tbeg, (int)(tend - tbeg), g.ibasis, g.B, sW, 1.0, sScan);
        for (int64_t t = tbeg + tid; t < tend; t += kGibbsThreads) inet[t] += delta * spk_u(spk, (int)(t - tbeg));
        return;
    }
    const XT* __restrict__ xcol = X + (int64_t)pre * g.B * g.T;
    for (int64_t t = tbeg + tid; t < tend; t += kGibbsThreads) {
        double u = 0.0;
        for (int b = 0; b < g.B; ++b) u += (double)xcol[(int64_t)b * g.T + t] * sW[b];
        inet[t] += delta * u;
    }
}

// I_net of the resident columns from the spikes alone (== seval(glm.I_net), glm.py:39): block = (chunk, column); the
// presynaptic neurons with A W != 0 are visited in index order and their currents accumulate in shared memory.
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_inet_spk_kernel(GibbsArgs g)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ int sScan[kGibbsThreads / 32 + 1];
    __shared__ double sW[kMaxBasis];
    const int nl = blockIdx.y, col = g.n_lo + nl;
    const int tid = threadIdx.x;
    const int64_t NB = (int64_t)g.N * g.B + g.F;
    const int64_t tbeg = (int64_t)blockIdx.x * kGibbsChunk;
    const int64_t tend = min(g.T, tbeg + kGibbsChunk);
    SpkSmem spk = spk_carve(dyn_smem, g.R);
    constexpr int kPer = kGibbsChunk / kGibbsThreads;      // bins per thread: tid, tid + 256, ...
    double acc[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) acc[j] = 0.0;
    const int nbins = (int)(tend - tbeg);
    for (int pre = 0; pre < g.N; ++pre) {
        const double aw = (double)g.A[(int64_t)pre * g.N + col] * g.W[(int64_t)pre * g.N + col];
        if (aw == 0.0) continue;                            // block-uniform
        __syncthreads();
        if (tid < kMaxBasis) sW[tid] = tid < g.B ? g.w[(int64_t)col * NB + (int64_t)pre * g.B + tid] : 0.0;
        __syncthreads();
        spk_build(spk, g.St + (int64_t)pre * g.ldst, g.halo, tbeg, nbins, g.ibasis, g.B, sW, aw, sScan);
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int i = tid + j * kGibbsThreads;
            if (i < nbins) acc[j] += spk_u(spk, i);
        }
    }
    double* __restrict__ inet = g.Inet + (int64_t)nl * g.T;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kGibbsThreads;
        if (i < nbins) inet[tbeg + i] = acc[j];
    }
}

int launch_gibbs_inet_from_spikes(const GibbsArgs& g, cudaStream_t stream)
{
    if (g.T <= 0 || g.ncols <= 0) return PYGLM_B200_OK;
    const size_t smem = spk_smem_bytes(g.R);
    PYGLM_CUDA(cudaFuncSetAttribute(gibbs_inet_spk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)g.nchunks, (unsigned)g.ncols);
    gibbs_inet_spk_kernel<<<grid, kGibbsThreads, smem, stream>>>(g);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

__global__ void gibbs_store_state_kernel(int8_t* A, double* W, int N, int M, const int32_t* cols,
                                         const int32_t* pres, const int8_t* anew, const double* wnew)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t idx = (int64_t)pres[m] * N + cols[m];
    A[idx] = anew[m];
    W[idx] = wnew[m];
}

// Xt[j][t] = X[t][j]: feature-major copy so that the B streams an edge reads are contiguous in time
template <typename XT>
__global__ void __launch_bounds__(256)
transpose_X_kernel(const XT* __restrict__ X, int64_t T, int64_t NB, int64_t ldx, XT* __restrict__ Xt)
{
    __shared__ XT tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int64_t j0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t t = t0 + r, j = j0 + tx;
        tile[r][tx] = (t < T && j < NB) ? X[t * ldx + j] : (XT)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = j0 + r, t = t0 + tx;
        if (j < NB && t < T) Xt[j * T + t] = tile[tx][r];
    }
}

int launch_transpose_X(const void* X, int64_t T, int64_t NB, int64_t ldx, int x_dtype, void* Xt, cudaStream_t stream)
{
    if (T <= 0) return PYGLM_B200_OK;
    dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(NB, 32));
    if (x_dtype == PYGLM_B200_X_F32)
        transpose_X_kernel<float><<<grid, 256, 0, stream>>>((const float*)X, T, NB, ldx, (float*)Xt);
    else
        transpose_X_kernel<double><<<grid, 256, 0, stream>>>((const double*)X, T, NB, ldx, (double*)Xt);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

int launch_gibbs_delta(const GibbsArgs& g, int M, const int32_t* d_cols, const int32_t* d_pres, int Q,
                       const double* d_wcand, double* d_out, cudaStream_t stream)
{
    if (M <= 0 || Q <= 0) return PYGLM_B200_OK;
    if (Q > kMaxCand) {
        set_error("gibbs_delta_ll: Q=%d > %d", Q, kMaxCand);
        return PYGLM_B200_EUNSUPPORTED;
    }
    dim3 grid((unsigned)g.nchunks, (unsigned)M);
    const bool f32 = g.x_dtype == PYGLM_B200_X_F32;
    const bool sp = g.nlin == PYGLM_B200_NLIN_SOFTPLUS;
    if (sp) { int rc = ensure_softplus_table(); if (rc) return rc; }
    if (g.spk) {                                            // u from the spikes: X is not read, one instantiation per (Q, nlin)
        const size_t smem = spk_smem_bytes(g.R);
#define PYGLM_GIBBS_SPK(QM)                                                                                            \
        do {                                                                                                           \
            auto kern = sp ? gibbs_delta_kernel<double, QM, PYGLM_B200_NLIN_SOFTPLUS, 1, true>                         \
                           : gibbs_delta_kernel<double, QM, PYGLM_B200_NLIN_EXP, 1, true>;                             \
            PYGLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
            kern<<<grid, kGibbsThreads, smem, stream>>>(g, d_cols, d_pres, Q, d_wcand);                                \
        } while (0)
        if (Q <= 4) PYGLM_GIBBS_SPK(4);
        else if (Q <= 11) PYGLM_GIBBS_SPK(11);
        else PYGLM_GIBBS_SPK(16);
#undef PYGLM_GIBBS_SPK
        PYGLM_CUDA(cudaGetLastError());
        gibbs_reduce_kernel<<<(unsigned)ceil_div((int64_t)M * Q, 128), 128, 0, stream>>>(g.partial, M, g.nchunks, Q, d_out);
        PYGLM_CUDA(cudaGetLastError());
        return PYGLM_B200_OK;
    }
#define PYGLM_GIBBS_LAUNCH3(QM, BM)                                                                                   \
    do {                                                                                                               \
        if (f32 && sp)       gibbs_delta_kernel<float, QM, PYGLM_B200_NLIN_SOFTPLUS, BM, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand); \
        else if (f32)        gibbs_delta_kernel<float, QM, PYGLM_B200_NLIN_EXP, BM, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);      \
        else if (sp)         gibbs_delta_kernel<double, QM, PYGLM_B200_NLIN_SOFTPLUS, BM, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand); \
        else                 gibbs_delta_kernel<double, QM, PYGLM_B200_NLIN_EXP, BM, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);     \
    } while (0)
#define PYGLM_GIBBS_LAUNCH(QM)                                                                                         \
    do {                                                                                                               \
        if (g.B <= 5) PYGLM_GIBBS_LAUNCH3(QM, 5);                                                                      \
        else if (g.B <= 10) PYGLM_GIBBS_LAUNCH3(QM, 10);                                                               \
        else PYGLM_GIBBS_LAUNCH3(QM, 16);                                                                              \
    } while (0)
    if (Q <= 4) PYGLM_GIBBS_LAUNCH(4);
    else if (Q <= 11) PYGLM_GIBBS_LAUNCH(11);
    else PYGLM_GIBBS_LAUNCH(16);
#undef PYGLM_GIBBS_LAUNCH3
#undef PYGLM_GIBBS_LAUNCH
    PYGLM_CUDA(cudaGetLastError());
    gibbs_reduce_kernel<<<(unsigned)ceil_div((int64_t)M * Q, 128), 128, 0, stream>>>(g.partial, M, g.nchunks, Q, d_out);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

int launch_gibbs_commit(const GibbsArgs& g, int M, const int32_t* d_cols, const int32_t* d_pres,
                        const int8_t* d_anew, const double* d_wnew, cudaStream_t stream)
{
    if (M <= 0) return PYGLM_B200_OK;
    dim3 grid((unsigned)g.nchunks, (unsigned)M);
    if (g.spk) {
        const size_t smem = spk_smem_bytes(g.R);
        PYGLM_CUDA(cudaFuncSetAttribute(gibbs_commit_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gibbs_commit_kernel<double, true><<<grid, kGibbsThreads, smem, stream>>>(g, d_cols, d_pres, d_anew, d_wnew);
    } else if (g.x_dtype == PYGLM_B200_X_F32)
        gibbs_commit_kernel<float, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, d_anew, d_wnew);
    else
        gibbs_commit_kernel<double, false><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, d_anew, d_wnew);
    PYGLM_CUDA(cudaGetLastError());
    gibbs_store_state_kernel<<<(unsigned)ceil_div(M, 128), 128, 0, stream>>>(g.A, g.W, g.N, M, d_cols, d_pres,
                                                                           d_anew, d_wnew);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

}  // namespace pyglm
