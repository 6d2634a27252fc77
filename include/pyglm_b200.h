/*
 * pyglm_b200.h -- C ABI of the B200-native engine for pyglm's data-parallel hot path.
 *
 * The reference (slinderman/theano_pyglm) has no FFI: its operator boundary is
 *   seval(expr, syms, vals)                       pyglm/utils/theano_func_wrapper.py:12-51
 * evaluated on Theano shared variables filled by
 *   Population.add_data / set_data                pyglm/population.py:198-231
 * Each entry point below replaces one family of `seval` call sites; the reference
 * file:line it stands in for is cited on the declaration.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - every function returns 0 on success or a negative PYGLM_B200_E* code;
 *     pyglm_b200_last_error() returns a thread-local message for the last failure.
 *   - numerics never raise: NaN / +-inf are produced exactly where the reference
 *     formula would produce them (callers keep their own guards,
 *     coord_descent.py:170-182, gibbs.py:1012-1019).
 *   - "host" entry points take host buffers and perform the H2D / D2H copies
 *     themselves (synchronous on return); "_dev" entry points take device pointers,
 *     enqueue on `stream` (a cudaStream_t passed as void*) and do not synchronise.
 *   - a dataset handle belongs to one GPU and one host thread at a time (the
 *     reference is single-threaded per engine, parallel_util.py:227).
 *   - neuron order, feature order (pre-major, basis fastest) and gradient order
 *     (bias, then w_ir) are the reference's sorted-key order
 *     (theano_func_wrapper.py:53-67, packvec.py:17-44).
 */
#ifndef PYGLM_B200_H
#define PYGLM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYGLM_B200_ABI_VERSION 5

#if defined(__GNUC__)
#define PYGLM_B200_API __attribute__((visibility("default")))
#else
#define PYGLM_B200_API
#endif

/* status codes */
#define PYGLM_B200_OK          0
#define PYGLM_B200_EINVAL     -1   /* bad shape / argument                       */
#define PYGLM_B200_ECUDA      -2   /* CUDA runtime error (see last_error)        */
#define PYGLM_B200_ENOMEM     -3   /* device allocation failed                   */
#define PYGLM_B200_ESTATE     -4   /* call out of order (e.g. gibbs before begin)*/
#define PYGLM_B200_EUNSUPPORTED -5 /* shape not supported by the requested path  */

/* nonlinearity: components/nlin.py:17-29 (exp) and :32-47 ('explinear' == softplus) */
#define PYGLM_B200_NLIN_EXP       0
#define PYGLM_B200_NLIN_SOFTPLUS  1

/* storage type of the filtered spike train X (reference: float64 `fS`) */
#define PYGLM_B200_X_F32  0
#define PYGLM_B200_X_F64  1
/* planes only: keep just the FP16 split planes of X that the tensor-core path streams (4 bytes per
 * element, written directly by the filter kernel); the FP64 path, firing_rate and get_fS are unavailable
 * (the Gibbs entry points gather their currents from the spikes).  For recordings whose FP32 X would not
 * fit next to the planes (configs C4/C5 shards). */
#define PYGLM_B200_X_PLANES 2
/* spikes only: neither X nor its planes are resident.  ll / gradient calls expand the spikes into the operand planes
 * chunk by chunk inside every evaluation (utils/basis.py:201-236 fused in front of glm.py:33-52; results bit-identical
 * to a planes-only dataset's up to the order of the last FP64 additions), and the Gibbs entry points gather the
 * presynaptic currents from the spike trains (one byte per bin instead of B filtered values).  This is what lets one GPU
 * of a neuron-sharded run -- the reference's own split, every engine with the whole data
 * (parallel_coord_descent.py:57-157, parallel_gibbs.py:162-168) -- hold ALL presynaptic data of a population whose X
 * would not fit (C4: 4 GB of spikes against 164 GB of X).  PATH_FP64, firing_rate and get_fS are unavailable. */
#define PYGLM_B200_X_NONE 3

/* arithmetic path for ll / gradient
 *   PATH_FP64  FP64 CUDA-core contractions over X as stored (exact path: 1e-11 on a float64 dataset).
 *   PATH_TC    tcgen05 `kind::f16` contractions on error-free FP16 splits of every operand
 *              (v*s = v1 + v2*2^-11, s a power of two: 22 significant bits, three products per contraction),
 *              FP32 accumulation per 128-bin tile in TMEM, FP32 epilogue (nonlinearity, Poisson terms), FP64 sums
 *              across tiles.  Holds 1e-6 (ll) / 1e-5 (gradient, max-norm) against the reference's float64.
 *   PATH_AUTO  PATH_TC where the dataset has it (FP32 X or planes), else PATH_FP64.  With the exp nonlinearity
 *              the epilogue flags every neuron whose activation leaves x <= 16 (e^x in FP32 no longer holds the
 *              tolerance there and overflows at 88); the host entry point re-evaluates flagged neurons on PATH_FP64
 *              before returning, the _dev entry point leaves the flags for pyglm_b200_range_flags(). */
#define PYGLM_B200_PATH_AUTO   0
#define PYGLM_B200_PATH_FP64   1
#define PYGLM_B200_PATH_TC     2

typedef struct pyglm_b200_dataset pyglm_b200_dataset;

PYGLM_B200_API const char* pyglm_b200_last_error(void);
PYGLM_B200_API int32_t     pyglm_b200_abi_version(void);

/* ------------------------------------------------------------------------------------
 * Data ingest + spike-history filtering (kernel K1).
 * Replaces LinearBasisImpulses.preprocess_data / DirichletImpulses.preprocess_data
 * (components/impulse.py:114-130, :378-394) -> convolve_with_basis
 * (utils/basis.py:201-236) and Glm.set_data / imp_model.set_data
 * (glm.py:99-110, impulse.py:132-133).
 *
 * S       host uint8 [(halo+T)][N] row-major spike counts (time-major like data["S"];
 *         counts are small non-negative integers, population.py:345-349).  The first
 *         `halo` rows are left context only (time-sharded use: the previous shard's
 *         last R bins); X and the likelihood cover rows [halo, halo+T).
 * ibasis  host float64 [R][B] interpolated basis (impulse.py:92-112); row k-1 is lag k.
 * X[t][pre*B+b] = sum_{k=1..R} ibasis[k-1][b] * S[t-k][pre]   (zero row prepended, basis.py:220)
 * ---------------------------------------------------------------------------------- */
PYGLM_B200_API int pyglm_b200_dataset_create(const uint8_t* S, int64_t T, int32_t halo, int32_t N, double dt,
                              const double* ibasis, int32_t R, int32_t B,
                              int32_t x_dtype, int32_t device,
                              pyglm_b200_dataset** out);

/* Same, with F = D_stim*B_stim filtered-stimulus features appended behind the N*B spike-history
 * features of X (F = 0, fstim = NULL: identical to dataset_create).
 * Replaces BasisStimulus.preprocess_data / set_data (components/bkgd.py:122-157): fstim is
 * data['fstim'], host float64 [T][F] (no halo rows), I_stim = fstim @ w_stim (bkgd.py:81).
 * With F > 0 every parameter row `w` and gradient row `out_g_w` below has N*B + F entries: the
 * N*B impulse weights followed by x['glms'][n]['bkgd']['w_stim']; stimulus features are not
 * masked by A / W.  Not available for PYGLM_B200_X_PLANES datasets. */
PYGLM_B200_API int pyglm_b200_dataset_create_stim(const uint8_t* S, int64_t T, int32_t halo, int32_t N, double dt,
                              const double* ibasis, int32_t R, int32_t B,
                              const double* fstim, int32_t F,
                              int32_t x_dtype, int32_t device,
                              pyglm_b200_dataset** out);
PYGLM_B200_API int pyglm_b200_dataset_destroy(pyglm_b200_dataset* ds);

/* number of stimulus features F of a dataset (0 for dataset_create) */
PYGLM_B200_API int32_t pyglm_b200_dataset_num_stim(const pyglm_b200_dataset* ds);

/* Dense causal basis projection of a real-valued signal (the stimulus):
 *   out[t][d*B+b] = sum_{k=1..R} ibasis[k-1][b] * stim[t-k][d]
 * == convolve_with_basis(stim, ibasis) flattened as in bkgd.py:145-154 (utils/basis.py:201-236).
 * stim host float64 [T][D], ibasis host float64 [R][B], out host float64 [T][D*B]. */
PYGLM_B200_API int pyglm_b200_filter_dense(const double* stim, int64_t T, int32_t D,
                            const double* ibasis, int32_t R, int32_t B,
                            int32_t device, double* out);

/* shape query: any pointer may be NULL */
PYGLM_B200_API int pyglm_b200_dataset_info(const pyglm_b200_dataset* ds, int64_t* T, int32_t* N, int32_t* B,
                            int32_t* R, int64_t* ldx, int32_t* x_dtype, int32_t* device);

/* copy the filtered spike train back: out is host float64 [T][N][B] == data['fS']
 * (impulse.py:130).  Parity / debugging only. */
PYGLM_B200_API int pyglm_b200_dataset_get_fS(const pyglm_b200_dataset* ds, double* out);

/* raw device pointers for zero-copy views (torch.from_blob style); owned by the handle */
PYGLM_B200_API void* pyglm_b200_dataset_device_X(const pyglm_b200_dataset* ds);
PYGLM_B200_API void* pyglm_b200_dataset_device_S(const pyglm_b200_dataset* ds);

/* re-run K1 on the resident spikes (bench / profiling of the filter alone): one pass rewrites everything K1 produced for
 * this dataset -- the filtered spike train X and, for FP32 / planes-only datasets without stimulus features, the FP16
 * split planes of the tensor-core path (utils/basis.py:201-236 fused with the operand split). */
PYGLM_B200_API int pyglm_b200_dataset_refilter(pyglm_b200_dataset* ds, void* stream);
/* bytes one such pass reads (spikes) and writes (X and / or planes): the algorithmic traffic of K1 */
PYGLM_B200_API int pyglm_b200_dataset_filter_bytes(const pyglm_b200_dataset* ds, int64_t* bytes_read, int64_t* bytes_written);

/* ------------------------------------------------------------------------------------
 * Log-likelihood and gradient for postsynaptic neurons n in [n_lo, n_hi)  (kernels K2f/K2b).
 * Replaces, for every n in the range, one call each of
 *   seval(glm.ll, syms, nvars)                    population.py:80-84, coord_descent.py:52-57
 *   seval(g_glm_ll, syms, nvars)                  coord_descent.py:30,72-77 (grads.py:9-28)
 * with glm.ll = sum_t(-dt*lam + log(lam)*S[:,n]) (glm.py:52),
 *      lam   = nlin(bias + I_imp @ (A[:,n]*W[:,n]))           (glm.py:33-45),
 *      I_imp = sum_b ir[t,pre,b] * w[pre,b]                   (impulse.py:58 / :308).
 *
 * bias  float64 [N]              x['glms'][n]['bias']['bias'][0]
 * w     float64 [N][N*B+F]       row n = x['glms'][n]['imp']['w_ir'] (pre-major, basis fastest), then
 *                                the F stimulus weights w_stim (dataset_create_stim; F = 0 otherwise);
 *                                for DirichletImpulses pass beta = |g|/sum|g| (impulse.py:286-291)
 * A     int8    [N][N] or NULL   x['net']['graph']['A'] (row = presynaptic); NULL = complete graph
 * W     float64 [N][N] or NULL   x['net']['weights']['W'] reshaped (N,N); NULL = ones (weights.py:32)
 * out_ll      float64 [n_hi-n_lo]
 * out_g_bias  float64 [n_hi-n_lo]          d ll_n / d bias_n            (may be NULL => ll only)
 * out_g_w     float64 [n_hi-n_lo][N*B+F]   d ll_n / d w[n][j]           (may be NULL => ll only)
 *
 * Host buffers may be pageable (they are staged through page-locked memory owned by the
 * handle) or page-locked (used in place).  From the second call with the same signature the
 * uploads, kernels and downloads replay as one CUDA graph.
 * Host and _dev entry points share the handle's workspaces: each use is ordered after the previous one with an
 * event, so the two may be mixed on one handle from one host thread without a device synchronise in between.
 * ---------------------------------------------------------------------------------- */
PYGLM_B200_API int pyglm_b200_ll_grad(pyglm_b200_dataset* ds,
                       const double* bias, const double* w, const int8_t* A, const double* W,
                       int32_t nlin, int32_t n_lo, int32_t n_hi, int32_t path,
                       double* out_ll, double* out_g_bias, double* out_g_w);

/* same, device pointers, asynchronous on `stream` */
PYGLM_B200_API int pyglm_b200_ll_grad_dev(pyglm_b200_dataset* ds,
                           const double* d_bias, const double* d_w, const int8_t* d_A, const double* d_W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi, int32_t path,
                           double* d_out_ll, double* d_out_g_bias, double* d_out_g_w,
                           void* stream);

/* which arithmetic path a call with `path` would take on this dataset: returns
 * PYGLM_B200_PATH_FP64 / PYGLM_B200_PATH_TC, or PYGLM_B200_EUNSUPPORTED. */
PYGLM_B200_API int pyglm_b200_resolve_path(const pyglm_b200_dataset* ds, int32_t path);

/* range flags of the last tensor-core evaluation: out_flags int32 [N], 1 = the neuron's activation left the range in
 * which the FP32 epilogue of the exp nonlinearity holds the tolerance (see PATH_AUTO).  Synchronises the device. */
PYGLM_B200_API int pyglm_b200_range_flags(const pyglm_b200_dataset* ds, int32_t* out_flags);

/* firing rate lam[t][n] for n in [n_lo,n_hi): host float64 [T][n_hi-n_lo].
 * Replaces seval(glm.lam, ...) in Population.eval_state (population.py:88-123); this is
 * the quantity the reference's only numeric assertion checks
 * (test/generate_synth_data.py:124-129). */
PYGLM_B200_API int pyglm_b200_firing_rate(pyglm_b200_dataset* ds,
                           const double* bias, const double* w, const int8_t* A, const double* W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi, double* out_lam);

/* ------------------------------------------------------------------------------------
 * Batched delta-log-likelihood for the collapsed Gibbs sampler over A/W (kernel K4).
 * Replaces CollapsedGibbsNetworkColumnUpdate._precompute_vars / _precompute_other_current /
 * _glm_ll (inference/gibbs.py:812-864, :910-937) for a batch of edges.
 *
 * gibbs_begin   uploads the state and makes I_net[:, n] resident for n in [n_lo, n_hi)
 *               (== seval(glm.I_net) with the current A, W; glm.py:39).
 * gibbs_delta_ll  for edge m = (pres[m] -> cols[m]) and candidate weights w_cand[m][0..Q):
 *               out_ll[m][q] = sum_t(-dt*f(x) + S[t,col] log f(x)),
 *               x = bias[col] + I_other[t] + w_cand[m][q] * I_imp[t, pre]     (gibbs.py:914)
 *               where I_other is I_net[:,col] with A[pre,col] := 0 (gibbs.py:838-861).
 *               Columns within one call must be distinct.
 * gibbs_commit  sets A[pre,col] = a_new, W[pre,col] = w_new and rank-1 updates the resident
 *               I_net (the reference instead recomputes the gemv for the next edge).
 * ---------------------------------------------------------------------------------- */
PYGLM_B200_API int pyglm_b200_gibbs_begin(pyglm_b200_dataset* ds,
                           const double* bias, const double* w, const int8_t* A, const double* W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi);
PYGLM_B200_API int pyglm_b200_gibbs_delta_ll(pyglm_b200_dataset* ds, int32_t M,
                              const int32_t* cols, const int32_t* pres,
                              int32_t Q, const double* w_cand, double* out_ll);
/* same, device pointers for cols / pres / w_cand / out_ll, asynchronous on `stream`; the edge list is
 * not validated (the caller guarantees distinct resident columns) */
PYGLM_B200_API int pyglm_b200_gibbs_delta_ll_dev(pyglm_b200_dataset* ds, int32_t M,
                              const int32_t* d_cols, const int32_t* d_pres,
                              int32_t Q, const double* d_w_cand, double* d_out_ll, void* stream);
PYGLM_B200_API int pyglm_b200_gibbs_commit(pyglm_b200_dataset* ds, int32_t M,
                            const int32_t* cols, const int32_t* pres,
                            const int8_t* a_new, const double* w_new);
/* copy the sampler's current A (int8 [N][N]) / W (float64 [N][N]) back; either may be NULL */
PYGLM_B200_API int pyglm_b200_gibbs_get_state(const pyglm_b200_dataset* ds, int8_t* A, double* W);
PYGLM_B200_API int pyglm_b200_gibbs_end(pyglm_b200_dataset* ds);

/* ------------------------------------------------------------------------------------
 * Sum over ranks of the time shards' partial results (one process per GPU).
 * Replaces the client-side sums of the reference's parallel drivers: ll / log_p summed over
 * engines (pyglm/utils/parallel_util.py:30,78) and over data sequences (population.py:41-43,
 * coord_descent.py:52-57) -- here [ll | g_bias | g_w] summed over the GPUs that each hold a
 * time shard of the recording.
 *
 * One-shot all-reduce over NVLink peer memory: every rank stores its vector into a slot of
 * every peer's receive buffer, raises a flag, and adds the slots in rank order (bitwise
 * identical results on all ranks, no host round trip).  Setup: comm_create on every rank,
 * comm_export -> 128 opaque bytes (two cudaIpc handles), exchange them out of band (e.g.
 * torch.distributed all_gather), comm_connect with the world*128 bytes in rank order.
 * allreduce_sum_dev is asynchronous on `stream`; d_in may equal d_out; every rank must call
 * it with the same n, in the same order.
 * ---------------------------------------------------------------------------------- */
typedef struct pyglm_b200_comm pyglm_b200_comm;
PYGLM_B200_API int pyglm_b200_comm_create(int32_t rank, int32_t world, int32_t device, int64_t max_doubles,
                           pyglm_b200_comm** out);
PYGLM_B200_API int pyglm_b200_comm_export(const pyglm_b200_comm* comm, void* handles128);
PYGLM_B200_API int pyglm_b200_comm_connect(pyglm_b200_comm* comm, const void* all_handles);
PYGLM_B200_API int pyglm_b200_allreduce_sum_dev(pyglm_b200_comm* comm, const double* d_in, double* d_out,
                                 int64_t n, void* stream);
PYGLM_B200_API int pyglm_b200_comm_destroy(pyglm_b200_comm* comm);
PYGLM_B200_API int32_t pyglm_b200_comm_world(const pyglm_b200_comm* comm);
/* Time-sharded evaluation in one call: ll / gradients of ALL N neurons on this rank's time shard, summed over the ranks of
 * `comm` (population.py:41-43: the reference adds the log-likelihoods of its data sequences the same way).  d_out is the
 * contiguous result vector [ll (N) | g_bias (N) | g_w (N x (N*B+F))] on the device, identical on every rank afterwards.
 * Where the fused kernel covers the population in one launch (N <= 32, 97..160 features) its final reduction IS the
 * collective (each block exchanges its row with the peers); otherwise the evaluation is followed by allreduce_sum_dev.
 * Asynchronous on `stream`; every rank must call it in the same order as its other collective calls on `comm`. */
PYGLM_B200_API int pyglm_b200_ll_grad_allreduce_dev(pyglm_b200_dataset* ds, pyglm_b200_comm* comm,
                                     const double* d_bias, const double* d_w, const int8_t* d_A, const double* d_W,
                                     int32_t nlin, int32_t path, double* d_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Measurement helper (bench.py): FP64 FMA throughput of this GPU's CUDA cores in TFLOP/s, from a register-resident
 * DFMA loop -- the roofline denominator of the Gibbs delta-ll kernel, which MEASURED_PEAKS.json does not carry.
 * ---------------------------------------------------------------------------------- */
PYGLM_B200_API int pyglm_b200_measure_fp64_peak(int32_t device, double* out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* PYGLM_B200_H */
