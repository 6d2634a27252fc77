// Shared declarations for the pyglm_b200 engine (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pyglm_b200.h"

namespace pyglm {

void set_error(const char* fmt, ...);

// Counts device (re)allocations of engine buffers.  A captured CUDA graph holds raw device pointers, so the host entry
// point re-captures whenever this has moved since its graph was built.
void note_allocation();
unsigned long long allocation_epoch();

#define PYGLM_CUDA(call)                                                               \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            ::pyglm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,           \
                               cudaGetErrorString(e_));                                \
            return (e_ == cudaErrorMemoryAllocation) ? PYGLM_B200_ENOMEM               \
                                                     : PYGLM_B200_ECUDA;               \
        }                                                                              \
    } while (0)

#define PYGLM_REQUIRE(cond, ...)                                                       \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            ::pyglm::set_error(__VA_ARGS__);                                           \
            return PYGLM_B200_EINVAL;                                                  \
        }                                                                              \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

constexpr int kMaxBasis = 16;   // B <= 16 (reference configs use 5 and 10)
constexpr int kMaxCand = 16;    // Q <= 16 candidate weights per edge (reference: 10 GH nodes + w=0)

// ---------------------------------------------------------------------------------
// Launchers (defined in the .cu files).  All enqueue on `stream` and return a status.
// ---------------------------------------------------------------------------------

// K1: X[t][pre*B+b] = sum_{k=1..R} ibasis[k-1][b] * S[halo+t-k][pre], written in one pass as the full-precision
// filtered spike train (X: FP32 or FP64, may be null) and / or as the FP16 split planes of the tensor-core path
// (X1, X2 with per-feature power-of-two scales sx; may be null): X * sx = X1 + X2 * 2^-11.
struct FilterOut {
    void* X = nullptr; int64_t ldx = 0; int x_dtype = PYGLM_B200_X_F32;
    __half* X1 = nullptr; __half* X2 = nullptr; int64_t ldp = 0; const float* sx = nullptr;
};
int launch_filter(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                  const FilterOut& out, cudaStream_t stream);

// dense FP64 causal filter of a real-valued stimulus (device pointers): out[t][d*B+b]
int launch_filter_dense(const double* d_stim, int64_t T, int D, const double* d_ibasis, int R, int B, double* d_out,
                        cudaStream_t stream);
// X[t][col0+f] = fstim[t][f]
int launch_fill_stim(const double* d_fstim, int64_t T, int F, void* dX, int64_t ldx, int64_t col0, int x_dtype,
                     cudaStream_t stream);

// St[n][t] = S[halo+t][n]
int launch_transpose_spikes(const uint8_t* dS, int64_t T, int N, int halo, uint8_t* dSt, cudaStream_t stream);

// Feature layout everywhere: NF = N*B + F columns; j < N*B is spike-history feature (pre = j/B, b = j%B),
// j >= N*B is stimulus feature j - N*B.  Parameter rows w[n][NF] and gradient rows g_w[n'][NF] share it.
// M[j][n'] = A[pre][n] W[pre][n] w[n][j]   (n = n_lo+n', pre = j/B; no mask for stimulus rows),
// zero padded to [NBp][Np];  Weff[n'][pre] = A[pre][n] W[pre][n]
int launch_build_M(const double* d_w, const int8_t* d_A, const double* d_W, int N, int B, int F, int n_lo, int ncols,
                   double* d_M, int Np, int64_t NBp, double* d_Weff, cudaStream_t stream);

struct SimtArgs {
    const void* X; int64_t ldx; int x_dtype;
    const uint8_t* S; int64_t T; int N; int halo; int B; int F;
    double dt; int nlin;
    int n_lo, ncols, Np;
    const double* bias;      // [N]
    const double* M;         // [NBp][Np]
    const double* Weff;      // [ncols][N]
    double* R;               // [T][Np] residuals (nullptr: forward only)
    double* llp; double* gbp;   // [tiles][Np] partials
    double* Gp; int splits;  // [splits][NBp64][Np]
    double* out_ll; double* out_gb; double* out_gw;   // [ncols], [ncols], [ncols][N*B+F]
    double* act_out;         // optional [ncols][T] activation without bias (Gibbs I_net)
    double* lam_out;         // optional [T][ncols] firing rate
};
int simt_workspace_tiles(int64_t T);
int simt_choose_splits(int64_t T, int64_t NF, int Np);
int launch_simt_ll_grad(const SimtArgs& a, cudaStream_t stream);

struct GibbsArgs {
    const void* X; int64_t ldx; int x_dtype;     // X here is the feature-major copy Xt[j][t] (unused when spk != 0)
    const uint8_t* St; int64_t ldst; int halo;   // spikes by column: St[n * ldst + t], t in [-halo, T)
    const double* ibasis; int R;                 // [R][B] interpolated basis (from-spikes mode)
    int spk;                                     // != 0: the presynaptic current u[t] is gathered from the spikes, X is never read
    int64_t T; int N; int B; int F;
    double dt; int nlin;
    int n_lo, ncols;
    const double* bias;      // [N]
    const double* w;         // [N][N*B+F]
    int8_t* A; double* W;    // device state [N][N]
    double* Inet;            // [ncols][T]
    double* partial;         // [M][nchunks][Q]
    int nchunks;
};
int gibbs_num_chunks(int64_t T);
int launch_transpose_X(const void* X, int64_t T, int64_t NB, int64_t ldx, int x_dtype, void* Xt, cudaStream_t stream);
int launch_gibbs_delta(const GibbsArgs& g, int M, const int32_t* d_cols, const int32_t* d_pres, int Q,
                       const double* d_wcand, double* d_out, cudaStream_t stream);
int launch_gibbs_commit(const GibbsArgs& g, int M, const int32_t* d_cols, const int32_t* d_pres,
                        const int8_t* d_anew, const double* d_wnew, cudaStream_t stream);
// from-spikes mode: I_net[nl][t] = sum_pre (A W)[pre][n] * sum_k h_pre[k] S[t-k][pre] for the resident columns
int launch_gibbs_inet_from_spikes(const GibbsArgs& g, cudaStream_t stream);
constexpr int kGibbsMaxLagsFromSpikes = 2048;

// ---------------------------------------------------------------------------------
// Device math shared by K2/K4: nonlinearity, its derivative and log, in FP64.
// components/nlin.py:25 (exp), :43/:47 (log(1+exp(x)) a.k.a. 'explinear').
// ---------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void nlin_eval(double x, int nlin, double& lam, double& dlam, double& loglam) {
    if (nlin == PYGLM_B200_NLIN_EXP) {
        lam = exp(x);
        dlam = lam;
        loglam = x;
    } else {
        const double e = exp(-fabs(x));
        const double l1p = log1p(e);
        lam = x > 0.0 ? x + l1p : l1p;
        dlam = x > 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
        loglam = log(lam);
    }
}
#endif

}  // namespace pyglm
