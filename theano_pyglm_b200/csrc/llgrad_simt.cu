// K2 (FP64 CUDA-core path): population log-likelihood and gradient.
//
//   x[t][n]   = bias[n] + sum_j X[t][j] M[j][n],  M[(pre,b)][n] = A[pre][n] W[pre][n] w[n][pre][b]
//                                                      glm.py:31-45, impulse.py:58
//   lam       = f(x)                                    nlin.py:25,43
//   ll[n]     = sum_t (-dt*lam + log(lam) * S[t][n])    glm.py:52
//   r[t][n]   = (S/lam - dt) * f'(x)                    == T.grad(glm.ll) through lam
//   g_bias[n] = sum_t r[t][n]
//   g_w[n][j] = A W [pre(j)][n] * sum_t X[t][j] r[t][n] coord_descent.py:30 / grads.py:9-28
//
// This is the exact-arithmetic path (every product and sum in FP64, deterministic
// reduction order); the tcgen05 path in llgrad_tc.cu is validated against it on device
// and both are validated against oracle/ in tests/.
#include "common.cuh"

namespace pyglm {

constexpr int kFwdTT = 64;       // time bins per forward tile
constexpr int kNT = 32;          // postsynaptic columns per tile
constexpr int kKC = 32;          // features per K step
constexpr int kThreads = 128;    // 16 x 8 threads, 4 x 4 outputs each
constexpr int kBwdJT = 64;       // features per backward tile
constexpr int kBwdTK = 32;       // time bins per backward K step

// ---------------------------------------------------------------------------------
__global__ void build_M_kernel(const double* __restrict__ w, const int8_t* __restrict__ A,
                               const double* __restrict__ W, int N, int B, int F, int n_lo, int ncols,
                               double* __restrict__ M, int Np, int64_t NBp, double* __restrict__ Weff)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t NB = (int64_t)N * B, NF = NB + F;
    if (idx < NBp * Np) {
        const int64_t j = idx / Np;
        const int nl = (int)(idx - j * Np);
        double v = 0.0;
        if (j < NF && nl < ncols) {
            const int n = n_lo + nl;
            if (j < NB) {
                const int pre = (int)(j / B);
                const double a = A ? (double)A[(int64_t)pre * N + n] : 1.0;
                const double ww = W ? W[(int64_t)pre * N + n] : 1.0;
                v = (a * ww) * w[(int64_t)n * NF + j];
            } else {
                v = w[(int64_t)n * NF + j];              // stimulus weight: no network mask (bkgd.py:81)
            }
        }
        M[idx] = v;
    }
    if (idx < (int64_t)ncols * N) {
        const int nl = (int)(idx / N);
        const int pre = (int)(idx - (int64_t)nl * N);
        const int n = n_lo + nl;
        const double a = A ? (double)A[(int64_t)pre * N + n] : 1.0;
        const double ww = W ? W[(int64_t)pre * N + n] : 1.0;
        Weff[idx] = a * ww;
    }
}

int launch_build_M(const double* d_w, const int8_t* d_A, const double* d_W, int N, int B, int F, int n_lo, int ncols,
                   double* d_M, int Np, int64_t NBp, double* d_Weff, cudaStream_t stream)
{
    const int64_t total = NBp * Np > (int64_t)ncols * N ? NBp * Np : (int64_t)ncols * N;
    build_M_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(d_w, d_A, d_W, N, B, F, n_lo, ncols,
                                                                        d_M, Np, NBp, d_Weff);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

// ---------------------------------------------------------------------------------
// Forward: activation tile (64 bins x 32 neurons) by a K loop over features, then the
// nonlinearity / Poisson term / residual epilogue and per-tile partial sums.
// ---------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(kThreads)
simt_fwd_kernel(SimtArgs a)
{
    __shared__ double sM[kKC][kNT];
    __shared__ XT sX[kFwdTT][kKC + 1];
    __shared__ double sRed[2][16][kNT];

    const XT* __restrict__ X = static_cast<const XT*>(a.X);
    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;            // 8 x 16
    const int64_t t0 = (int64_t)blockIdx.x * kFwdTT;
    const int nt0 = blockIdx.y * kNT;
    const int64_t NB = (int64_t)a.N * a.B + a.F;     // all features: spike history + stimulus

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int64_t k0 = 0; k0 < NB; k0 += kKC) {
        for (int e = tid; e < kFwdTT * kKC; e += kThreads) {
            const int r = e / kKC, kk = e - r * kKC;
            const int64_t t = t0 + r, k = k0 + kk;
            sX[r][kk] = (t < a.T && k < NB) ? X[t * a.ldx + k] : (XT)0;
        }
        for (int e = tid; e < kKC * kNT; e += kThreads) {
            const int kk = e / kNT, c = e - kk * kNT;
            const int64_t k = k0 + kk;
            sM[kk][c] = (k < NB) ? a.M[k * a.Np + nt0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kKC; ++kk) {
            double xv[4], mv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = (double)sX[4 * ty + i][kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) mv[j] = sM[kk][4 * tx + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(xv[i], mv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
    double pll[4], pgb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { pll[j] = 0.0; pgb[j] = 0.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t t = t0 + 4 * ty + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nl = nt0 + 4 * tx + j;
            double rr = 0.0;
            if (t < a.T && nl < a.ncols) {
                const int n = a.n_lo + nl;
                const double x = a.bias[n] + acc[i][j];
                double lam, dlam, loglam;
                nlin_eval(x, a.nlin, lam, dlam, loglam);
                const double s = (double)a.S[((int64_t)a.halo + t) * a.N + n];
                pll[j] += -a.dt * lam + loglam * s;
                rr = (s / lam - a.dt) * dlam;
                pgb[j] += rr;
                if (a.act_out) a.act_out[(int64_t)nl * a.T + t] = acc[i][j];
                if (a.lam_out) a.lam_out[t * a.ncols + nl] = lam;
            }
            if (a.R && t < a.T) a.R[t * a.Np + nl] = rr;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sRed[0][ty][4 * tx + j] = pll[j];
        sRed[1][ty][4 * tx + j] = pgb[j];
    }
    __syncthreads();
    if (tid < 2 * kNT) {
        const int which = tid / kNT, c = tid - which * kNT;
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += sRed[which][r][c];
        double* dst = which ? a.gbp : a.llp;
        dst[(int64_t)blockIdx.x * a.Np + nt0 + c] = s;
    }
}

// Sum the per-tile partials in a fixed order: one block per (column, quantity).
__global__ void __launch_bounds__(256)
reduce_tiles_kernel(const double* __restrict__ llp, const double* __restrict__ gbp, int ntiles, int Np, int ncols,
                    double* __restrict__ out_ll, double* __restrict__ out_gb)
{
    __shared__ double sh[256];
    const int nl = blockIdx.x;
    const double* src = blockIdx.y ? gbp : llp;
    double* dst = blockIdx.y ? out_gb : out_ll;
    if (dst == nullptr) return;
    double s = 0.0;
    for (int i = threadIdx.x; i < ntiles; i += 256) s += src[(int64_t)i * Np + nl];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && nl < ncols) dst[nl] = sh[0];
}

// ---------------------------------------------------------------------------------
// Backward: G[j][n] = sum_t X[t][j] r[t][n], split over time; partials reduced in order.
// ---------------------------------------------------------------------------------
template <typename XT>
__global__ void __launch_bounds__(kThreads)
simt_bwd_kernel(SimtArgs a, int64_t chunk, int64_t NBp64)
{
    __shared__ __align__(16) XT sX[kBwdTK][kBwdJT];
    __shared__ __align__(16) double sR[kBwdTK][kNT];

    const XT* __restrict__ X = static_cast<const XT*>(a.X);
    const int tid = threadIdx.x;
    const int tx = tid & 7, ty = tid >> 3;            // n cols 4tx.., features 4ty..
    const int64_t j0 = (int64_t)blockIdx.x * kBwdJT;
    const int nt0 = blockIdx.y * kNT;
    const int64_t tbeg = (int64_t)blockIdx.z * chunk;
    const int64_t tend = min(a.T, tbeg + chunk);
    const int64_t NB = (int64_t)a.N * a.B + a.F;     // all features: spike history + stimulus

    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int64_t tk = tbeg; tk < tend; tk += kBwdTK) {
        for (int e = tid; e < kBwdTK * kBwdJT; e += kThreads) {
            const int r = e / kBwdJT, c = e - r * kBwdJT;
            const int64_t t = tk + r, j = j0 + c;
            sX[r][c] = (t < tend && j < NB) ? X[t * a.ldx + j] : (XT)0;
        }
        for (int e = tid; e < kBwdTK * kNT; e += kThreads) {
            const int r = e / kNT, c = e - r * kNT;
            const int64_t t = tk + r;
            sR[r][c] = (t < tend) ? a.R[t * a.Np + nt0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kBwdTK; ++r) {
            double xv[4], rv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = (double)sX[r][4 * ty + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) rv[j] = sR[r][4 * tx + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(xv[i], rv[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* gp = a.Gp + ((int64_t)blockIdx.z * NBp64 + j0) * a.Np + nt0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gp[(int64_t)(4 * ty + i) * a.Np + 4 * tx + j] = acc[i][j];
}

__global__ void __launch_bounds__(256)
reduce_G_kernel(const double* __restrict__ Gp, int splits, int64_t NBp64, int Np, int N, int B, int F, int ncols,
                const double* __restrict__ Weff, double* __restrict__ out_gw)
{
    const int64_t NB = (int64_t)N * B + F;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NB * ncols) return;
    const int nl = (int)(idx / NB);
    const int64_t j = idx - (int64_t)nl * NB;
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += Gp[((int64_t)z * NBp64 + j) * Np + nl];
    out_gw[idx] = j < (int64_t)N * B ? Weff[(int64_t)nl * N + (j / B)] * s : s;
}

int simt_workspace_tiles(int64_t T) { return (int)ceil_div(T, kFwdTT); }

int simt_choose_splits(int64_t T, int64_t NF, int Np)
{
    const int64_t tiles = ceil_div(NF, kBwdJT) * (Np / kNT);
    int64_t splits = ceil_div(148 * 8, tiles);
    const int64_t max_splits = ceil_div(T, 4 * kBwdTK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 1024) splits = 1024;
    return (int)splits;
}

template <typename XT>
static int launch_simt_t(const SimtArgs& a, cudaStream_t stream)
{
    const int ntiles = simt_workspace_tiles(a.T);
    dim3 gridf((unsigned)ntiles, (unsigned)(a.Np / kNT));
    simt_fwd_kernel<XT><<<gridf, kThreads, 0, stream>>>(a);
    PYGLM_CUDA(cudaGetLastError());
    dim3 gridr((unsigned)a.Np, 2);
    reduce_tiles_kernel<<<gridr, 256, 0, stream>>>(a.llp, a.gbp, ntiles, a.Np, a.ncols, a.out_ll, a.out_gb);
    PYGLM_CUDA(cudaGetLastError());
    if (a.R && a.out_gw) {
        const int64_t NB = (int64_t)a.N * a.B + a.F;     // all features: spike history + stimulus
        const int64_t NBp64 = round_up(NB, kBwdJT);
        int64_t chunk = round_up(ceil_div(a.T, a.splits), kBwdTK);
        dim3 gridb((unsigned)(NBp64 / kBwdJT), (unsigned)(a.Np / kNT), (unsigned)a.splits);
        simt_bwd_kernel<XT><<<gridb, kThreads, 0, stream>>>(a, chunk, NBp64);
        PYGLM_CUDA(cudaGetLastError());
        reduce_G_kernel<<<(unsigned)ceil_div(NB * a.ncols, 256), 256, 0, stream>>>(
            a.Gp, a.splits, NBp64, a.Np, a.N, a.B, a.F, a.ncols, a.Weff, a.out_gw);
        PYGLM_CUDA(cudaGetLastError());
    }
    return PYGLM_B200_OK;
}

int launch_simt_ll_grad(const SimtArgs& a, cudaStream_t stream)
{
    if (a.T <= 0 || a.ncols <= 0) return PYGLM_B200_OK;
    return a.x_dtype == PYGLM_B200_X_F32 ? launch_simt_t<float>(a, stream) : launch_simt_t<double>(a, stream);
}

}  // namespace pyglm
