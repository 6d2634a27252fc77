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
#include "common.cuh"

namespace pyglm {

constexpr int kGibbsThreads = 256;
constexpr int kGibbsChunk = 8192;     // bins per block

int gibbs_num_chunks(int64_t T) { return (int)ceil_div(T, kGibbsChunk); }

template <typename XT, int QMAX>
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_delta_kernel(GibbsArgs g, const int32_t* __restrict__ cols, const int32_t* __restrict__ pres,
                   int Q, const double* __restrict__ wcand)
{
    __shared__ double sW[kMaxBasis];
    __shared__ double sRed[QMAX][kGibbsThreads / 32];

    const XT* __restrict__ X = static_cast<const XT*>(g.X);
    const int m = blockIdx.y;
    const int col = cols[m], pre = pres[m];
    const int nl = col - g.n_lo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t NB = (int64_t)g.N * g.B;

    if (tid < g.B) sW[tid] = g.w[(int64_t)col * NB + (int64_t)pre * g.B + tid];
    __syncthreads();

    const double aw_old = (double)g.A[(int64_t)pre * g.N + col] * g.W[(int64_t)pre * g.N + col];
    const double bias = g.bias[col];
    double wq[QMAX], acc[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        wq[q] = (q < Q) ? wcand[(int64_t)m * Q + q] : 0.0;
        acc[q] = 0.0;
    }

    const int64_t tbeg = (int64_t)blockIdx.x * kGibbsChunk;
    const int64_t tend = min(g.T, tbeg + kGibbsChunk);
    const double* __restrict__ inet = g.Inet + (int64_t)nl * g.T;
    const uint8_t* __restrict__ st = g.St + (int64_t)col * g.T;
    const XT* __restrict__ xcol = X + (int64_t)pre * g.B;

    for (int64_t t = tbeg + tid; t < tend; t += kGibbsThreads) {
        const XT* xr = xcol + t * g.ldx;
        double u = 0.0;
        for (int b = 0; b < g.B; ++b) u += (double)xr[b] * sW[b];
        const double base = inet[t] - aw_old * u;
        const double s = (double)st[t];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            if (q < Q) {
                const double x = bias + (base + wq[q] * u);
                double lam, dlam, loglam;
                nlin_eval(x, g.nlin, lam, dlam, loglam);
                acc[q] += -g.dt * lam + loglam * s;
            }
        }
    }
    // block reduction, fixed order: lanes (xor tree), then warps 0..7
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        double v = acc[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sRed[q][warp] = v;
    }
    __syncthreads();
    if (tid < Q) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kGibbsThreads / 32; ++w) v += sRed[tid][w];
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

template <typename XT>
__global__ void __launch_bounds__(kGibbsThreads)
gibbs_commit_kernel(GibbsArgs g, const int32_t* __restrict__ cols, const int32_t* __restrict__ pres,
                    const int8_t* __restrict__ anew, const double* __restrict__ wnew)
{
    __shared__ double sW[kMaxBasis];
    const XT* __restrict__ X = static_cast<const XT*>(g.X);
    const int m = blockIdx.y;
    const int col = cols[m], pre = pres[m];
    const int nl = col - g.n_lo;
    const int tid = threadIdx.x;
    const int64_t NB = (int64_t)g.N * g.B;
    if (tid < g.B) sW[tid] = g.w[(int64_t)col * NB + (int64_t)pre * g.B + tid];
    __syncthreads();
    const double aw_old = (double)g.A[(int64_t)pre * g.N + col] * g.W[(int64_t)pre * g.N + col];
    const double aw_new = (double)anew[m] * wnew[m];
    const double delta = aw_new - aw_old;
    if (delta == 0.0) return;
    const int64_t tbeg = (int64_t)blockIdx.x * kGibbsChunk;
    const int64_t tend = min(g.T, tbeg + kGibbsChunk);
    double* __restrict__ inet = g.Inet + (int64_t)nl * g.T;
    const XT* __restrict__ xcol = X + (int64_t)pre * g.B;
    for (int64_t t = tbeg + tid; t < tend; t += kGibbsThreads) {
        const XT* xr = xcol + t * g.ldx;
        double u = 0.0;
        for (int b = 0; b < g.B; ++b) u += (double)xr[b] * sW[b];
        inet[t] += delta * u;
    }
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
    if (Q <= 1) {
        if (f32) gibbs_delta_kernel<float, 1><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
        else     gibbs_delta_kernel<double, 1><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
    } else if (Q <= 4) {
        if (f32) gibbs_delta_kernel<float, 4><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
        else     gibbs_delta_kernel<double, 4><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
    } else if (Q <= 11) {
        if (f32) gibbs_delta_kernel<float, 11><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
        else     gibbs_delta_kernel<double, 11><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
    } else {
        if (f32) gibbs_delta_kernel<float, 16><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
        else     gibbs_delta_kernel<double, 16><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, Q, d_wcand);
    }
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
    if (g.x_dtype == PYGLM_B200_X_F32)
        gibbs_commit_kernel<float><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, d_anew, d_wnew);
    else
        gibbs_commit_kernel<double><<<grid, kGibbsThreads, 0, stream>>>(g, d_cols, d_pres, d_anew, d_wnew);
    PYGLM_CUDA(cudaGetLastError());
    gibbs_store_state_kernel<<<(unsigned)ceil_div(M, 128), 128, 0, stream>>>(g.A, g.W, g.N, M, d_cols, d_pres,
                                                                           d_anew, d_wnew);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

}  // namespace pyglm
