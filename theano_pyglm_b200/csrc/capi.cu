// C ABI of the engine (include/pyglm_b200.h): dataset handle, host<->device plumbing, dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <new>
#include <vector>

#include "allreduce.cuh"
#include "common.cuh"
#include "llgrad_tc.cuh"

namespace pyglm {

static thread_local char g_err[1024] = "";

static std::atomic<unsigned long long> g_alloc_epoch{0};
void note_allocation() { g_alloc_epoch.fetch_add(1, std::memory_order_relaxed); }
unsigned long long allocation_epoch() { return g_alloc_epoch.load(std::memory_order_relaxed); }

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t count)
    {
        if (count <= n) return PYGLM_B200_OK;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        note_allocation();
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return PYGLM_B200_ENOMEM;
        }
        n = count;
        return PYGLM_B200_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace pyglm

using namespace pyglm;

struct pyglm_b200_dataset {
    int device = 0;
    int64_t T = 0;
    int N = 0, B = 0, R = 0, halo = 0;
    int F = 0;                        // stimulus features behind the N*B spike-history features of X
    int64_t NF() const { return (int64_t)N * B + F; }
    double dt = 0.0;
    int x_dtype = PYGLM_B200_X_F32;
    int64_t ldx = 0;
    cudaStream_t stream = nullptr;

    DevBuf<uint8_t> S, St;
    DevBuf<unsigned char> X, Xt;      // Xt: feature-major copy of X, built at the first gibbs_begin
    bool xt_ready = false;
    DevBuf<double> ibasis;

    // parameter staging (host entry points) and workspaces
    DevBuf<double> p_bias, p_w, p_W, M, Weff, Rres, llp, gbp, Gp, o_ll, o_gb, o_gw, lam;
    DevBuf<int8_t> p_A;
    TcWorkspace tc;

    // host-buffer entry point (pyglm_b200_ll_grad): pinned staging + a CUDA graph of the whole call, replayed while the
    // call signature stays the same (an optimiser calls it hundreds of times with new parameter values only)
    struct HostCall {
        double *bias = nullptr, *w = nullptr, *W = nullptr, *ll = nullptr, *gb = nullptr, *gw = nullptr;
        int8_t* A = nullptr;
        unsigned* flags = nullptr;                    // [N] range flags of the last call (always page-locked)
        bool staging_ready = false;                   // all seven staging buffers exist
        long long key[16] = {-1};                     // nlin, n_lo, n_hi, path, hasA, hasW, want_gb, want_gw, direct, 7 pointers
        int seen = 0;                                 // calls with this key so far
        cudaGraphExec_t exec = nullptr;
        unsigned long long epoch = 0;                 // allocation_epoch() when `exec` was captured
        bool disabled = false;                        // capture failed once: plain stream calls from then on
    } hc;

    // The workspaces above are shared by every entry point.  The host entry points run on `stream`, the _dev ones on
    // the caller's: `ws_done` is recorded after each use and the next user's stream waits on it when it is another one.
    cudaEvent_t ws_done = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;

    // Gibbs state
    bool gibbs_active = false;
    bool g_spk = false;                // from-spikes mode: currents gathered from St, X never read
    int g_nlo = 0, g_ncols = 0, g_nlin = 0;
    DevBuf<double> g_bias, g_w, g_W, Inet, partial, g_wcand, g_out, g_wnew;
    DevBuf<int8_t> g_A, g_anew;
    DevBuf<int32_t> g_cols, g_pres;
};

#define DS_GUARD(ds)                                                   \
    PYGLM_REQUIRE((ds) != nullptr, "null dataset handle");             \
    PYGLM_CUDA(cudaSetDevice((ds)->device))

#define TRY(expr)                                                      \
    do { int rc_ = (expr); if (rc_ != PYGLM_B200_OK) return rc_; } while (0)

// order this use of the handle's workspaces after the previous one when that ran on another stream
static int ws_acquire(pyglm_b200_dataset* ds, cudaStream_t st)
{
    if (ds->ws_used && ds->ws_stream != st) PYGLM_CUDA(cudaStreamWaitEvent(st, ds->ws_done, 0));
    return PYGLM_B200_OK;
}
static int ws_release(pyglm_b200_dataset* ds, cudaStream_t st)
{
    PYGLM_CUDA(cudaEventRecord(ds->ws_done, st));
    ds->ws_stream = st;
    ds->ws_used = true;
    return PYGLM_B200_OK;
}

extern "C" {

const char* pyglm_b200_last_error(void) { return g_err; }
int32_t pyglm_b200_abi_version(void) { return PYGLM_B200_ABI_VERSION; }

int pyglm_b200_dataset_create(const uint8_t* S, int64_t T, int32_t halo, int32_t N, double dt,
                              const double* ibasis, int32_t R, int32_t B,
                              int32_t x_dtype, int32_t device, pyglm_b200_dataset** out)
{
    return pyglm_b200_dataset_create_stim(S, T, halo, N, dt, ibasis, R, B, nullptr, 0, x_dtype, device, out);
}

int32_t pyglm_b200_dataset_num_stim(const pyglm_b200_dataset* ds) { return ds ? ds->F : 0; }

int pyglm_b200_filter_dense(const double* stim, int64_t T, int32_t D, const double* ibasis, int32_t R, int32_t B,
                            int32_t device, double* out)
{
    PYGLM_REQUIRE(T >= 0 && D >= 1, "filter_dense: bad shape T=%lld D=%d", (long long)T, D);
    PYGLM_REQUIRE(R >= 1 && B >= 1 && B <= kMaxBasis, "filter_dense: bad basis shape R=%d B=%d (B<=%d)", R, B, kMaxBasis);
    PYGLM_REQUIRE(ibasis != nullptr && (T == 0 || (stim != nullptr && out != nullptr)), "filter_dense: null argument");
    if (T == 0) return PYGLM_B200_OK;
    PYGLM_CUDA(cudaSetDevice(device));
    DevBuf<double> d_stim, d_ib, d_out;
    int rc;
    if ((rc = d_stim.ensure((size_t)T * D)) || (rc = d_ib.ensure((size_t)R * B)) || (rc = d_out.ensure((size_t)T * D * B))) {
        d_stim.release(); d_ib.release(); d_out.release();
        return rc;
    }
    auto done = [&](int r) { d_stim.release(); d_ib.release(); d_out.release(); return r; };
    cudaError_t e;
    if ((e = cudaMemcpy(d_stim.p, stim, (size_t)T * D * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(d_ib.p, ibasis, (size_t)R * B * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
        set_error("filter_dense: upload failed: %s", cudaGetErrorString(e));
        return done(PYGLM_B200_ECUDA);
    }
    if ((rc = launch_filter_dense(d_stim.p, T, D, d_ib.p, R, B, d_out.p, nullptr))) return done(rc);
    if ((e = cudaMemcpy(out, d_out.p, (size_t)T * D * B * sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess) {
        set_error("filter_dense: download failed: %s", cudaGetErrorString(e));
        return done(PYGLM_B200_ECUDA);
    }
    return done(PYGLM_B200_OK);
}

int pyglm_b200_dataset_create_stim(const uint8_t* S, int64_t T, int32_t halo, int32_t N, double dt,
                                   const double* ibasis, int32_t R, int32_t B,
                                   const double* fstim, int32_t F,
                                   int32_t x_dtype, int32_t device, pyglm_b200_dataset** out)
{
    PYGLM_REQUIRE(out != nullptr, "dataset_create: out is null");
    *out = nullptr;
    PYGLM_REQUIRE(T >= 0 && N >= 1 && halo >= 0, "dataset_create: bad shape T=%lld N=%d halo=%d", (long long)T, N, halo);
    PYGLM_REQUIRE(R >= 1 && B >= 1 && B <= kMaxBasis, "dataset_create: bad basis shape R=%d B=%d (B<=%d)", R, B, kMaxBasis);
    PYGLM_REQUIRE(S != nullptr || (T + halo) == 0, "dataset_create: S is null");
    PYGLM_REQUIRE(ibasis != nullptr, "dataset_create: ibasis is null");
    PYGLM_REQUIRE(x_dtype == PYGLM_B200_X_F32 || x_dtype == PYGLM_B200_X_F64 || x_dtype == PYGLM_B200_X_PLANES ||
                  x_dtype == PYGLM_B200_X_NONE, "dataset_create: bad x_dtype %d", x_dtype);
    PYGLM_REQUIRE(dt > 0.0, "dataset_create: dt must be positive");
    PYGLM_REQUIRE(F >= 0 && (F == 0 || fstim != nullptr || T == 0), "dataset_create: bad stimulus block F=%d", F);
    if (F > 0 && (x_dtype == PYGLM_B200_X_PLANES || x_dtype == PYGLM_B200_X_NONE)) {
        set_error("dataset_create: stimulus features need a resident filtered spike train (x_dtype F32 / F64)");
        return PYGLM_B200_EUNSUPPORTED;
    }
    PYGLM_CUDA(cudaSetDevice(device));

    pyglm_b200_dataset* ds = new (std::nothrow) pyglm_b200_dataset();
    PYGLM_REQUIRE(ds != nullptr, "dataset_create: host allocation failed");
    ds->device = device; ds->T = T; ds->N = N; ds->B = B; ds->R = R; ds->halo = halo; ds->dt = dt;
    ds->x_dtype = x_dtype; ds->F = F;
    ds->ldx = round_up(ds->NF(), 4);
    auto fail = [&](int rc) { pyglm_b200_dataset_destroy(ds); return rc; };

    cudaError_t e = cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return fail(PYGLM_B200_ECUDA); }
    e = cudaEventCreateWithFlags(&ds->ws_done, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("cudaEventCreate: %s", cudaGetErrorString(e)); return fail(PYGLM_B200_ECUDA); }

    const size_t nS = (size_t)(T + halo) * N;
    const size_t esz = x_dtype == PYGLM_B200_X_F64 ? 8 : 4;
    const bool planes_only = x_dtype == PYGLM_B200_X_PLANES || x_dtype == PYGLM_B200_X_NONE;   // no FP32 / FP64 X resident
    int rc;
    if ((rc = ds->S.ensure(nS ? nS : 1)) || (rc = ds->St.ensure(nS ? nS : 1)) ||
        (rc = ds->X.ensure(planes_only || (size_t)T * ds->ldx * esz == 0 ? 1 : (size_t)T * ds->ldx * esz)) ||
        (rc = ds->ibasis.ensure((size_t)R * B)))
        return fail(rc);
#define CK(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { set_error("%s -> %s", #call, cudaGetErrorString(e2)); return fail(PYGLM_B200_ECUDA); } } while (0)
    if (nS) CK(cudaMemcpyAsync(ds->S.p, S, nS, cudaMemcpyHostToDevice, ds->stream));
    CK(cudaMemcpyAsync(ds->ibasis.p, ibasis, (size_t)R * B * sizeof(double), cudaMemcpyHostToDevice, ds->stream));
    CK(cudaMemsetAsync(ds->X.p, 0, ds->X.n, ds->stream));
    if (x_dtype == PYGLM_B200_X_NONE) {
        // spikes only: nothing resident to filter.  The Gibbs entry points gather their currents from the spikes; ll / gradient
        // evaluations expand the spikes into the operand planes chunk by chunk (from-spikes K2, llgrad_tc.cu)
        if (T > 0 && (rc = tc_prepare_streamed(ds->tc, ds->S.p, T, N, halo, ds->ibasis.p, R, B, ds->stream))) return fail(rc);
    } else if (planes_only) {
        if (T > 0 && (rc = tc_build_planes_direct(ds->tc, ds->S.p, T, N, halo, ds->ibasis.p, R, B, nullptr, 0, ds->stream))) return fail(rc);
    } else {
        // single-pass ingest: an FP32 dataset without stimulus features gets its tensor-core planes from the same K1 pass
        bool done = false;
        const char* lazy = getenv("PYGLM_LAZY_PLANES");       // tests: build the planes from the resident X at the first tensor-core call
        if (x_dtype == PYGLM_B200_X_F32 && F == 0 && T > 0 && tc_supported(T, N, B, x_dtype) && !(lazy && atoi(lazy))) {
            rc = tc_build_planes_direct(ds->tc, ds->S.p, T, N, halo, ds->ibasis.p, R, B, (float*)ds->X.p, ds->ldx, ds->stream);
            if (rc == PYGLM_B200_OK) done = true;
            else if (rc == PYGLM_B200_ENOMEM) { ds->tc.release(); cudaGetLastError(); }   // no room for the planes now: X alone
            else return fail(rc);
        }
        if (!done) {
            FilterOut fo;
            fo.X = ds->X.p; fo.ldx = ds->ldx; fo.x_dtype = x_dtype;
            if ((rc = launch_filter(ds->S.p, T, N, halo, ds->ibasis.p, R, B, fo, ds->stream))) return fail(rc);
        }
    }
    // spikes by column, halo bins included: St[n][halo + T] (K4 reads per-column streams, and in from-spikes mode the
    // R bins before a chunk)
    if ((rc = launch_transpose_spikes(ds->S.p, T + halo, N, 0, ds->St.p, ds->stream))) return fail(rc);
    if (F > 0 && T > 0) {
        DevBuf<double> d_fs;
        if ((rc = d_fs.ensure((size_t)T * F))) return fail(rc);
        cudaError_t e3 = cudaMemcpyAsync(d_fs.p, fstim, (size_t)T * F * sizeof(double), cudaMemcpyHostToDevice, ds->stream);
        if (e3 == cudaSuccess) rc = launch_fill_stim(d_fs.p, T, F, ds->X.p, ds->ldx, (int64_t)N * B, x_dtype, ds->stream);
        cudaStreamSynchronize(ds->stream);
        d_fs.release();
        if (e3 != cudaSuccess) { set_error("dataset_create: stimulus upload failed: %s", cudaGetErrorString(e3)); return fail(PYGLM_B200_ECUDA); }
        if (rc) return fail(rc);
    }
    CK(cudaStreamSynchronize(ds->stream));
#undef CK
    *out = ds;
    return PYGLM_B200_OK;
}

int pyglm_b200_dataset_destroy(pyglm_b200_dataset* ds)
{
    if (!ds) return PYGLM_B200_OK;
    cudaSetDevice(ds->device);
    if (ds->stream) { cudaStreamSynchronize(ds->stream); }
    ds->S.release(); ds->St.release(); ds->X.release(); ds->Xt.release(); ds->ibasis.release();
    ds->p_bias.release(); ds->p_w.release(); ds->p_W.release(); ds->p_A.release();
    ds->M.release(); ds->Weff.release(); ds->Rres.release(); ds->llp.release(); ds->gbp.release();
    ds->Gp.release(); ds->o_ll.release(); ds->o_gb.release(); ds->o_gw.release(); ds->lam.release();
    ds->g_bias.release(); ds->g_w.release(); ds->g_W.release(); ds->g_A.release(); ds->Inet.release();
    ds->partial.release(); ds->g_wcand.release(); ds->g_out.release(); ds->g_wnew.release();
    ds->g_anew.release(); ds->g_cols.release(); ds->g_pres.release();
    ds->tc.release();
    if (ds->hc.exec) cudaGraphExecDestroy(ds->hc.exec);
    cudaFreeHost(ds->hc.bias); cudaFreeHost(ds->hc.w); cudaFreeHost(ds->hc.W); cudaFreeHost(ds->hc.A);
    cudaFreeHost(ds->hc.ll); cudaFreeHost(ds->hc.gb); cudaFreeHost(ds->hc.gw); cudaFreeHost(ds->hc.flags);
    if (ds->ws_done) cudaEventDestroy(ds->ws_done);
    if (ds->stream) cudaStreamDestroy(ds->stream);
    delete ds;
    return PYGLM_B200_OK;
}

int pyglm_b200_dataset_info(const pyglm_b200_dataset* ds, int64_t* T, int32_t* N, int32_t* B,
                            int32_t* R, int64_t* ldx, int32_t* x_dtype, int32_t* device)
{
    PYGLM_REQUIRE(ds != nullptr, "null dataset handle");
    if (T) *T = ds->T;
    if (N) *N = ds->N;
    if (B) *B = ds->B;
    if (R) *R = ds->R;
    if (ldx) *ldx = ds->ldx;
    if (x_dtype) *x_dtype = ds->x_dtype;
    if (device) *device = ds->device;
    return PYGLM_B200_OK;
}

void* pyglm_b200_dataset_device_X(const pyglm_b200_dataset* ds) { return ds ? (void*)ds->X.p : nullptr; }
void* pyglm_b200_dataset_device_S(const pyglm_b200_dataset* ds) { return ds ? (void*)ds->S.p : nullptr; }

int pyglm_b200_dataset_get_fS(const pyglm_b200_dataset* ds, double* out)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(out != nullptr, "get_fS: out is null");
    if (ds->x_dtype == PYGLM_B200_X_PLANES || ds->x_dtype == PYGLM_B200_X_NONE) { set_error("get_fS: this dataset keeps no filtered spike train"); return PYGLM_B200_EUNSUPPORTED; }
    const int64_t NB = (int64_t)ds->N * ds->B;
    if (ds->T == 0) return PYGLM_B200_OK;
    PYGLM_CUDA(cudaStreamSynchronize(ds->stream));
    if (ds->x_dtype == PYGLM_B200_X_F64) {
        PYGLM_CUDA(cudaMemcpy2D(out, NB * sizeof(double), ds->X.p, ds->ldx * sizeof(double),
                                NB * sizeof(double), ds->T, cudaMemcpyDeviceToHost));
    } else {
        // stream back in slabs and widen on the host
        const int64_t slab = 1 << 16;
        std::vector<float> tmp((size_t)slab * NB);
        for (int64_t t0 = 0; t0 < ds->T; t0 += slab) {
            const int64_t nt = (ds->T - t0 < slab) ? ds->T - t0 : slab;
            PYGLM_CUDA(cudaMemcpy2D(tmp.data(), NB * sizeof(float),
                                    (const float*)ds->X.p + t0 * ds->ldx, ds->ldx * sizeof(float),
                                    NB * sizeof(float), nt, cudaMemcpyDeviceToHost));
            double* o = out + t0 * NB;
            for (int64_t i = 0; i < nt * NB; ++i) o[i] = (double)tmp[i];
        }
    }
    return PYGLM_B200_OK;
}

// what one K1 pass over this dataset writes: the resident X and / or the split planes (spike-history features only)
static FilterOut filter_outputs(pyglm_b200_dataset* ds)
{
    FilterOut fo;
    if (ds->x_dtype == PYGLM_B200_X_F32 || ds->x_dtype == PYGLM_B200_X_F64) { fo.X = ds->X.p; fo.ldx = ds->ldx; fo.x_dtype = ds->x_dtype; }
    if (ds->x_dtype != PYGLM_B200_X_F64 && ds->F == 0 && ds->tc.planes_ready) {
        fo.X1 = ds->tc.X1; fo.X2 = ds->tc.X2; fo.ldp = ds->tc.ldp; fo.sx = ds->tc.sx;
    }
    return fo;
}

int pyglm_b200_dataset_refilter(pyglm_b200_dataset* ds, void* stream)
{
    DS_GUARD(ds);
    if (ds->x_dtype == PYGLM_B200_X_NONE) { set_error("refilter: this dataset keeps no filtered spike train"); return PYGLM_B200_EUNSUPPORTED; }
    ds->xt_ready = false;
    return launch_filter(ds->S.p, ds->T, ds->N, ds->halo, ds->ibasis.p, ds->R, ds->B, filter_outputs(ds), (cudaStream_t)stream);
}

int pyglm_b200_dataset_filter_bytes(const pyglm_b200_dataset* ds, int64_t* bytes_read, int64_t* bytes_written)
{
    PYGLM_REQUIRE(ds != nullptr, "null dataset handle");
    const FilterOut fo = filter_outputs(const_cast<pyglm_b200_dataset*>(ds));
    if (bytes_read) *bytes_read = (ds->T + ds->halo) * ds->N;
    if (bytes_written)
        *bytes_written = (fo.X ? ds->T * (int64_t)ds->N * ds->B * (fo.x_dtype == PYGLM_B200_X_F64 ? 8 : 4) : 0) +
                         (fo.X1 ? ds->T * fo.ldp * 4 : 0);
    return PYGLM_B200_OK;
}

// ------------------------------------------------------------------------------------
static int resolve_path(const pyglm_b200_dataset* ds, int path, bool need_aux)
{
    if (ds->x_dtype == PYGLM_B200_X_NONE) return (path != PYGLM_B200_PATH_FP64 && !need_aux && ds->tc.streamed) ? PYGLM_B200_PATH_TC : -1;
    const bool planes_only = ds->x_dtype == PYGLM_B200_X_PLANES;
    if (path == PYGLM_B200_PATH_FP64) return planes_only ? -1 : PYGLM_B200_PATH_FP64;
    const bool tc_ok = (planes_only || tc_supported(ds->T, ds->N, ds->B, ds->x_dtype)) && !need_aux;
    if (planes_only) return tc_ok ? PYGLM_B200_PATH_TC : -1;
    if (path == PYGLM_B200_PATH_TC) return tc_ok ? PYGLM_B200_PATH_TC : -1;
    return tc_ok ? PYGLM_B200_PATH_TC : PYGLM_B200_PATH_FP64;
}

static int ll_grad_dev_impl(pyglm_b200_dataset* ds,
                            const double* d_bias, const double* d_w, const int8_t* d_A, const double* d_W,
                            int nlin, int n_lo, int n_hi, int path,
                            double* d_ll, double* d_gb, double* d_gw,
                            double* d_act, double* d_lam, cudaStream_t stream, const void* ar = nullptr)
{
    PYGLM_REQUIRE(nlin == PYGLM_B200_NLIN_EXP || nlin == PYGLM_B200_NLIN_SOFTPLUS, "bad nlin %d", nlin);
    PYGLM_REQUIRE(0 <= n_lo && n_lo <= n_hi && n_hi <= ds->N, "bad neuron range [%d,%d) for N=%d", n_lo, n_hi, ds->N);
    PYGLM_REQUIRE(d_bias && d_w && d_ll, "ll_grad: null argument");
    PYGLM_REQUIRE((d_gb == nullptr) == (d_gw == nullptr), "ll_grad: pass both gradient outputs or neither");
    const int ncols = n_hi - n_lo;
    if (ncols == 0) return PYGLM_B200_OK;
    const int64_t NB = ds->NF();
    if (ds->T == 0) {   // empty recording: ll = 0, gradients = 0
        PYGLM_CUDA(cudaMemsetAsync(d_ll, 0, ncols * sizeof(double), stream));
        if (d_gb) PYGLM_CUDA(cudaMemsetAsync(d_gb, 0, ncols * sizeof(double), stream));
        if (d_gw) PYGLM_CUDA(cudaMemsetAsync(d_gw, 0, ncols * NB * sizeof(double), stream));
        return PYGLM_B200_OK;
    }
    const int use = resolve_path(ds, path, d_act != nullptr || d_lam != nullptr);
    if (use < 0) {
        set_error("ll_grad: requested path is not available for this dataset (N=%d B=%d x_dtype=%d)", ds->N, ds->B, ds->x_dtype);
        return PYGLM_B200_EUNSUPPORTED;
    }
    if (use == PYGLM_B200_PATH_TC) {
        TcArgs t{};
        t.X = (ds->x_dtype == PYGLM_B200_X_PLANES || ds->x_dtype == PYGLM_B200_X_NONE) ? nullptr : (const float*)ds->X.p; t.ldx = ds->ldx; t.S = ds->S.p; t.T = ds->T; t.N = ds->N; t.halo = ds->halo;
        t.B = ds->B; t.F = ds->F; t.ibasis = ds->ibasis.p; t.R = ds->R; t.dt = ds->dt; t.nlin = nlin; t.n_lo = n_lo; t.ncols = ncols;
        t.bias = d_bias; t.w = d_w; t.A = d_A; t.W = d_W;
        t.out_ll = d_ll; t.out_gb = d_gb; t.out_gw = d_gw;
        t.flags = nullptr;
        t.ar = ar;
        if (nlin == PYGLM_B200_NLIN_EXP) {          // range flags of this call's columns (kExpSafe, tc_common.cuh)
            if (!ds->tc.planes_ready) { int rc = tc_ensure_planes(t, ds->tc, stream); if (rc) return rc; }
            PYGLM_CUDA(cudaMemsetAsync(ds->tc.colflag + n_lo, 0, (size_t)ncols * sizeof(unsigned), stream));
            t.flags = ds->tc.colflag;
        }
        return ds->tc.streamed ? launch_tc_ll_grad_streamed(t, ds->tc, stream) : launch_tc_ll_grad(t, ds->tc, stream);
    }

    const int Np = (int)round_up(ncols, 32);
    const int64_t NBp = NB;
    TRY(ds->M.ensure((size_t)NBp * Np));
    TRY(ds->Weff.ensure((size_t)ncols * ds->N));
    const int ntiles = simt_workspace_tiles(ds->T);
    TRY(ds->llp.ensure((size_t)ntiles * Np));
    TRY(ds->gbp.ensure((size_t)ntiles * Np));
    SimtArgs a{};
    a.X = ds->X.p; a.ldx = ds->ldx; a.x_dtype = ds->x_dtype;
    a.S = ds->S.p; a.T = ds->T; a.N = ds->N; a.halo = ds->halo; a.B = ds->B; a.F = ds->F;
    a.dt = ds->dt; a.nlin = nlin; a.n_lo = n_lo; a.ncols = ncols; a.Np = Np;
    a.bias = d_bias; a.M = ds->M.p; a.Weff = ds->Weff.p;
    a.llp = ds->llp.p; a.gbp = ds->gbp.p;
    a.out_ll = d_ll; a.out_gb = d_gb; a.out_gw = d_gw;
    a.act_out = d_act; a.lam_out = d_lam;
    if (d_gw) {
        TRY(ds->Rres.ensure((size_t)ds->T * Np));
        a.R = ds->Rres.p;
        a.splits = simt_choose_splits(ds->T, NB, Np);
        TRY(ds->Gp.ensure((size_t)a.splits * round_up(NB, 64) * Np));
        a.Gp = ds->Gp.p;
    }
    TRY(launch_build_M(d_w, d_A, d_W, ds->N, ds->B, ds->F, n_lo, ncols, ds->M.p, Np, NBp, ds->Weff.p, stream));
    return launch_simt_ll_grad(a, stream);
}

int pyglm_b200_resolve_path(const pyglm_b200_dataset* ds, int32_t path)
{
    PYGLM_REQUIRE(ds != nullptr, "null dataset handle");
    const int use = resolve_path(ds, path, false);
    return use < 0 ? PYGLM_B200_EUNSUPPORTED : use;
}

int pyglm_b200_ll_grad_dev(pyglm_b200_dataset* ds,
                           const double* d_bias, const double* d_w, const int8_t* d_A, const double* d_W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi, int32_t path,
                           double* d_out_ll, double* d_out_g_bias, double* d_out_g_w, void* stream)
{
    DS_GUARD(ds);
    TRY(ws_acquire(ds, (cudaStream_t)stream));
    TRY(ll_grad_dev_impl(ds, d_bias, d_w, d_A, d_W, nlin, n_lo, n_hi, path, d_out_ll, d_out_g_bias, d_out_g_w,
                         nullptr, nullptr, (cudaStream_t)stream));
    return ws_release(ds, (cudaStream_t)stream);
}

int pyglm_b200_ll_grad_allreduce_dev(pyglm_b200_dataset* ds, pyglm_b200_comm* comm,
                                     const double* d_bias, const double* d_w, const int8_t* d_A, const double* d_W,
                                     int32_t nlin, int32_t path, double* d_out, void* stream)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(d_out != nullptr, "ll_grad_allreduce: null output");
    const int N = ds->N;
    const int64_t NB = ds->NF(), n = (int64_t)N * (2 + NB);
    const int world = comm ? pyglm_b200_comm_world(comm) : 1;
    cudaStream_t st = (cudaStream_t)stream;
    TRY(ws_acquire(ds, st));
    const bool fuse = world > 1 && ds->T > 0 && !ds->tc.streamed && resolve_path(ds, path, false) == PYGLM_B200_PATH_TC &&
                      tc_can_fuse_allreduce(N, NB);
    if (fuse) {                                              // the evaluation's final reduction is the collective
        ArEpoch ep;
        TRY(ar_begin_epoch(comm, n, &ep));
        TRY(ll_grad_dev_impl(ds, d_bias, d_w, d_A, d_W, nlin, 0, N, path, d_out, d_out + N, d_out + 2 * (int64_t)N, nullptr, nullptr, st, &ep));
    } else {
        TRY(ll_grad_dev_impl(ds, d_bias, d_w, d_A, d_W, nlin, 0, N, path, d_out, d_out + N, d_out + 2 * (int64_t)N, nullptr, nullptr, st));
        if (world > 1) TRY(pyglm_b200_allreduce_sum_dev(comm, d_out, d_out, n, stream));
    }
    return ws_release(ds, st);
}

int pyglm_b200_range_flags(const pyglm_b200_dataset* ds, int32_t* out_flags)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(out_flags != nullptr, "range_flags: out is null");
    for (int n = 0; n < ds->N; ++n) out_flags[n] = 0;
    if (!ds->tc.colflag) return PYGLM_B200_OK;
    PYGLM_CUDA(cudaDeviceSynchronize());
    PYGLM_CUDA(cudaMemcpy(out_flags, ds->tc.colflag, (size_t)ds->N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return PYGLM_B200_OK;
}

// upload the parameter block of a host call into the handle's staging buffers
static int stage_params(pyglm_b200_dataset* ds, const double* bias, const double* w, const int8_t* A, const double* W,
                        DevBuf<double>& b_bias, DevBuf<double>& b_w, DevBuf<int8_t>& b_A, DevBuf<double>& b_W,
                        cudaStream_t stream)
{
    PYGLM_REQUIRE(bias && w, "null bias / w");
    const size_t N = ds->N, NB = (size_t)ds->NF();
    TRY(b_bias.ensure(N));
    TRY(b_w.ensure(N * NB));
    PYGLM_CUDA(cudaMemcpyAsync(b_bias.p, bias, N * sizeof(double), cudaMemcpyHostToDevice, stream));
    PYGLM_CUDA(cudaMemcpyAsync(b_w.p, w, N * NB * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (A) {
        TRY(b_A.ensure(N * N));
        PYGLM_CUDA(cudaMemcpyAsync(b_A.p, A, N * N, cudaMemcpyHostToDevice, stream));
    }
    if (W) {
        TRY(b_W.ensure(N * N));
        PYGLM_CUDA(cudaMemcpyAsync(b_W.p, W, N * N * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    return PYGLM_B200_OK;
}

// everything one host-buffer call enqueues: parameter upload from pinned host memory, the evaluation, result download
struct HostPtrs {
    const double* bias; const double* w; const int8_t* A; const double* W;
    double* ll; double* gb; double* gw;
    unsigned* flags;
    bool mapped_out = true;       // outputs are page-locked host memory the kernels may write directly
};

static int enqueue_host_call(pyglm_b200_dataset* ds, const HostPtrs& hp, int nlin, int n_lo, int n_hi, int path, cudaStream_t st)
{
    const size_t N = ds->N, NF = (size_t)ds->NF();
    const int ncols = n_hi - n_lo;
    const bool grad = hp.gb || hp.gw;
    PYGLM_CUDA(cudaMemcpyAsync(ds->p_bias.p, hp.bias, N * sizeof(double), cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->p_w.p, hp.w, N * NF * sizeof(double), cudaMemcpyHostToDevice, st));
    if (hp.A) PYGLM_CUDA(cudaMemcpyAsync(ds->p_A.p, hp.A, N * N, cudaMemcpyHostToDevice, st));
    if (hp.W) PYGLM_CUDA(cudaMemcpyAsync(ds->p_W.p, hp.W, N * N * sizeof(double), cudaMemcpyHostToDevice, st));
    // Results: the final-reduction kernels write them straight into the caller's page-locked buffers when those are mapped
    // into the device's address space (three fewer copy nodes per call: ~4 % of the C2 call).  The from-spikes path
    // accumulates into its outputs (reads them back), so it keeps device buffers and copies.
    double *d_ll = nullptr, *d_gb = nullptr, *d_gw = nullptr;
    bool zero_copy = hp.mapped_out && !ds->tc.streamed && !(ds->T == 0);
    if (zero_copy) {
        zero_copy = cudaHostGetDevicePointer((void**)&d_ll, hp.ll, 0) == cudaSuccess &&
                    (!hp.gb || cudaHostGetDevicePointer((void**)&d_gb, hp.gb, 0) == cudaSuccess) &&
                    (!hp.gw || cudaHostGetDevicePointer((void**)&d_gw, hp.gw, 0) == cudaSuccess);
        if (!zero_copy) cudaGetLastError();
    }
    if (zero_copy) {
        TRY(ll_grad_dev_impl(ds, ds->p_bias.p, ds->p_w.p, hp.A ? ds->p_A.p : nullptr, hp.W ? ds->p_W.p : nullptr,
                             nlin, n_lo, n_hi, path, d_ll, grad ? d_gb : nullptr, grad ? d_gw : nullptr, nullptr, nullptr, st));
    } else {
    TRY(ll_grad_dev_impl(ds, ds->p_bias.p, ds->p_w.p, hp.A ? ds->p_A.p : nullptr, hp.W ? ds->p_W.p : nullptr,
                         nlin, n_lo, n_hi, path, ds->o_ll.p, grad ? ds->o_gb.p : nullptr, grad ? ds->o_gw.p : nullptr,
                         nullptr, nullptr, st));
    PYGLM_CUDA(cudaMemcpyAsync(hp.ll, ds->o_ll.p, ncols * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (hp.gb) PYGLM_CUDA(cudaMemcpyAsync(hp.gb, ds->o_gb.p, ncols * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (hp.gw) PYGLM_CUDA(cudaMemcpyAsync(hp.gw, ds->o_gw.p, (size_t)ncols * NF * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (hp.flags)   // range flags of the FP32 epilogues (exp nonlinearity on the tensor-core path)
        PYGLM_CUDA(cudaMemcpyAsync(hp.flags + n_lo, ds->tc.colflag + n_lo, ncols * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    return PYGLM_B200_OK;
}

static bool is_pinned_host(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int pyglm_b200_ll_grad(pyglm_b200_dataset* ds,
                       const double* bias, const double* w, const int8_t* A, const double* W,
                       int32_t nlin, int32_t n_lo, int32_t n_hi, int32_t path,
                       double* out_ll, double* out_g_bias, double* out_g_w)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(out_ll != nullptr, "ll_grad: out_ll is null");
    PYGLM_REQUIRE(bias && w, "null bias / w");
    PYGLM_REQUIRE(0 <= n_lo && n_lo <= n_hi && n_hi <= ds->N, "bad neuron range [%d,%d) for N=%d", n_lo, n_hi, ds->N);
    const int ncols = n_hi - n_lo;
    if (ncols == 0) return PYGLM_B200_OK;
    const size_t N = ds->N, NF = (size_t)ds->NF();
    cudaStream_t st = ds->stream;
    auto& h = ds->hc;
    const bool want_gb = out_g_bias != nullptr, want_gw = out_g_w != nullptr, grad = want_gb || want_gw;

    TRY(ds->p_bias.ensure(N)); TRY(ds->p_w.ensure(N * NF));
    if (A) TRY(ds->p_A.ensure(N * N));
    if (W) TRY(ds->p_W.ensure(N * N));
    TRY(ds->o_ll.ensure(ds->N));
    if (grad) { TRY(ds->o_gb.ensure(ds->N)); TRY(ds->o_gw.ensure(N * NF)); }

    // Caller buffers that are page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are used as they
    // are; pageable ones go through the handle's own pinned staging.
    const bool direct = is_pinned_host(bias) && is_pinned_host(w) && is_pinned_host(A) && is_pinned_host(W) &&
                        is_pinned_host(out_ll) && is_pinned_host(out_g_bias) && is_pinned_host(out_g_w);
    // PATH_AUTO with the exp nonlinearity on the tensor-core path: the epilogue flags every column whose activation
    // leaves the range in which FP32 e^x holds the tolerance; those columns are re-evaluated on the FP64 path below
    const bool heal = path == PYGLM_B200_PATH_AUTO && nlin == PYGLM_B200_NLIN_EXP && ds->T > 0 &&
                      ds->x_dtype != PYGLM_B200_X_PLANES && ds->x_dtype != PYGLM_B200_X_NONE &&
                      resolve_path(ds, path, false) == PYGLM_B200_PATH_TC;
    if (heal && !h.flags) {
        PYGLM_CUDA(cudaMallocHost(&h.flags, N * sizeof(unsigned)));
        memset(h.flags, 0, N * sizeof(unsigned));
    }
    HostPtrs hp{bias, w, A, W, out_ll, out_g_bias, out_g_w, heal ? h.flags : nullptr};
    if (!direct) {
        if (!h.staging_ready) {     // all or nothing: a failed allocation must not leave a half-built set behind
            void** bufs[7] = {(void**)&h.bias, (void**)&h.ll, (void**)&h.gb, (void**)&h.W, (void**)&h.A, (void**)&h.w, (void**)&h.gw};
            const size_t sizes[7] = {N * sizeof(double), N * sizeof(double), N * sizeof(double), N * N * sizeof(double), N * N,
                                     N * NF * sizeof(double), N * NF * sizeof(double)};
            for (int i = 0; i < 7; ++i) {
                cudaError_t e = cudaMallocHost(bufs[i], sizes[i]);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    for (int j = 0; j < 7; ++j) { if (*bufs[j]) cudaFreeHost(*bufs[j]); *bufs[j] = nullptr; }
                    set_error("ll_grad: page-locked staging allocation (%zu bytes) failed: %s", sizes[i], cudaGetErrorString(e));
                    return PYGLM_B200_ENOMEM;
                }
            }
            h.staging_ready = true;
        }
        memcpy(h.bias, bias, N * sizeof(double));
        memcpy(h.w, w, N * NF * sizeof(double));
        if (A) memcpy(h.A, A, N * N);
        if (W) memcpy(h.W, W, N * N * sizeof(double));
        hp = HostPtrs{h.bias, h.w, A ? h.A : nullptr, W ? h.W : nullptr, h.ll, want_gb ? h.gb : nullptr, want_gw ? h.gw : nullptr,
                      heal ? h.flags : nullptr};
    }

    // the whole call is a CUDA graph from the second call with the same signature (and, for caller-pinned buffers,
    // the same addresses) on
    const long long key[16] = {nlin, n_lo, n_hi, path, A != nullptr, W != nullptr, want_gb, want_gw, direct,
                               (long long)(uintptr_t)hp.bias, (long long)(uintptr_t)hp.w, (long long)(uintptr_t)hp.A,
                               (long long)(uintptr_t)hp.W, (long long)(uintptr_t)hp.ll, (long long)(uintptr_t)hp.gb,
                               (long long)(uintptr_t)hp.gw};
    TRY(ws_acquire(ds, st));
    if (memcmp(key, h.key, sizeof(key)) != 0) {
        if (h.exec) { cudaGraphExecDestroy(h.exec); h.exec = nullptr; }
        memcpy(h.key, key, sizeof(key));
        h.seen = 0;
    }
    h.seen += 1;
    if (h.exec && h.epoch != allocation_epoch()) {       // some engine buffer moved since the capture: pointers may be stale
        cudaGraphExecDestroy(h.exec);
        h.exec = nullptr;
        h.seen = 1;                                      // run once plainly (lets every buffer settle), then capture again
    }
    if (h.exec) {
        PYGLM_CUDA(cudaGraphLaunch(h.exec, st));
    } else if (h.seen >= 2 && !h.disabled && ds->T > 0) {
        // second call with this signature: every buffer it needs exists by now, so the enqueue is pure stream work
        cudaGraph_t graph = nullptr;
        int rc = PYGLM_B200_OK;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = enqueue_host_call(ds, hp, nlin, n_lo, n_hi, path, st);
            e = cudaStreamEndCapture(st, &graph);
        }
        if (e == cudaSuccess && rc == PYGLM_B200_OK && graph) e = cudaGraphInstantiate(&h.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess || rc != PYGLM_B200_OK || !h.exec) {
            cudaGetLastError();
            h.exec = nullptr;
            h.disabled = true;
            TRY(enqueue_host_call(ds, hp, nlin, n_lo, n_hi, path, st));
        } else {
            h.epoch = allocation_epoch();
            PYGLM_CUDA(cudaGraphLaunch(h.exec, st));
        }
    } else {
        TRY(enqueue_host_call(ds, hp, nlin, n_lo, n_hi, path, st));
    }
    TRY(ws_release(ds, st));
    PYGLM_CUDA(cudaStreamSynchronize(st));
    if (heal) {
        for (int lo = n_lo; lo < n_hi;) {
            if (!h.flags[lo]) { ++lo; continue; }
            int hi = lo;
            while (hi < n_hi && h.flags[hi]) ++hi;          // a run of flagged columns: one FP64 evaluation
            const int off = lo - n_lo, nrun = hi - lo;
            TRY(ll_grad_dev_impl(ds, ds->p_bias.p, ds->p_w.p, hp.A ? ds->p_A.p : nullptr, hp.W ? ds->p_W.p : nullptr,
                                 nlin, lo, hi, PYGLM_B200_PATH_FP64, ds->o_ll.p + off, grad ? ds->o_gb.p + off : nullptr,
                                 grad ? ds->o_gw.p + (size_t)off * NF : nullptr, nullptr, nullptr, st));
            PYGLM_CUDA(cudaMemcpyAsync(hp.ll + off, ds->o_ll.p + off, nrun * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (hp.gb) PYGLM_CUDA(cudaMemcpyAsync(hp.gb + off, ds->o_gb.p + off, nrun * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (hp.gw) PYGLM_CUDA(cudaMemcpyAsync(hp.gw + (size_t)off * NF, ds->o_gw.p + (size_t)off * NF,
                                                  (size_t)nrun * NF * sizeof(double), cudaMemcpyDeviceToHost, st));
            lo = hi;
        }
        TRY(ws_release(ds, st));
        PYGLM_CUDA(cudaStreamSynchronize(st));
    }
    if (!direct) {
        memcpy(out_ll, h.ll, ncols * sizeof(double));
        if (want_gb) memcpy(out_g_bias, h.gb, ncols * sizeof(double));
        if (want_gw) memcpy(out_g_w, h.gw, (size_t)ncols * NF * sizeof(double));
    }
    return PYGLM_B200_OK;
}

int pyglm_b200_firing_rate(pyglm_b200_dataset* ds,
                           const double* bias, const double* w, const int8_t* A, const double* W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi, double* out_lam)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(out_lam != nullptr, "firing_rate: out is null");
    PYGLM_REQUIRE(0 <= n_lo && n_lo <= n_hi && n_hi <= ds->N, "bad neuron range [%d,%d) for N=%d", n_lo, n_hi, ds->N);
    const int ncols = n_hi - n_lo;
    if (ncols == 0 || ds->T == 0) return PYGLM_B200_OK;
    cudaStream_t st = ds->stream;
    TRY(stage_params(ds, bias, w, A, W, ds->p_bias, ds->p_w, ds->p_A, ds->p_W, st));
    TRY(ds->o_ll.ensure(ncols));
    TRY(ds->lam.ensure((size_t)ds->T * ncols));
    TRY(ll_grad_dev_impl(ds, ds->p_bias.p, ds->p_w.p, A ? ds->p_A.p : nullptr, W ? ds->p_W.p : nullptr,
                         nlin, n_lo, n_hi, PYGLM_B200_PATH_FP64, ds->o_ll.p, nullptr, nullptr,
                         nullptr, ds->lam.p, st));
    PYGLM_CUDA(cudaMemcpyAsync(out_lam, ds->lam.p, (size_t)ds->T * ncols * sizeof(double), cudaMemcpyDeviceToHost, st));
    PYGLM_CUDA(cudaStreamSynchronize(st));
    return PYGLM_B200_OK;
}

// ------------------------------------------------------------------------------------
// Gibbs
// ------------------------------------------------------------------------------------
static GibbsArgs gibbs_args(pyglm_b200_dataset* ds)
{
    GibbsArgs g{};
    g.X = ds->Xt.p; g.ldx = ds->ldx; g.x_dtype = ds->x_dtype;
    g.ldst = ds->T + ds->halo; g.halo = ds->halo; g.St = ds->St.p + ds->halo;
    g.ibasis = ds->ibasis.p; g.R = ds->R; g.spk = ds->g_spk ? 1 : 0;
    g.T = ds->T; g.N = ds->N; g.B = ds->B; g.F = ds->F;
    g.dt = ds->dt; g.nlin = ds->g_nlin; g.n_lo = ds->g_nlo; g.ncols = ds->g_ncols;
    g.bias = ds->g_bias.p; g.w = ds->g_w.p; g.A = ds->g_A.p; g.W = ds->g_W.p;
    g.Inet = ds->Inet.p; g.partial = ds->partial.p; g.nchunks = gibbs_num_chunks(ds->T);
    return g;
}

int pyglm_b200_gibbs_begin(pyglm_b200_dataset* ds,
                           const double* bias, const double* w, const int8_t* A, const double* W,
                           int32_t nlin, int32_t n_lo, int32_t n_hi)
{
    DS_GUARD(ds);
    PYGLM_REQUIRE(nlin == PYGLM_B200_NLIN_EXP || nlin == PYGLM_B200_NLIN_SOFTPLUS, "bad nlin %d", nlin);
    PYGLM_REQUIRE(0 <= n_lo && n_lo < n_hi && n_hi <= ds->N, "bad neuron range [%d,%d) for N=%d", n_lo, n_hi, ds->N);
    PYGLM_REQUIRE(A && W, "gibbs_begin needs explicit A and W");
    // From-spikes mode (forced by PYGLM_GIBBS_FROM_SPIKES=1, the only mode of planes-only / spikes-only datasets): the
    // presynaptic currents are gathered from the spike trains, no feature-major copy of X is built or read.
    bool spk = ds->x_dtype == PYGLM_B200_X_PLANES || ds->x_dtype == PYGLM_B200_X_NONE;
    if (const char* env = getenv("PYGLM_GIBBS_FROM_SPIKES")) spk = spk || atoi(env) != 0;
    if (spk && (ds->F > 0 || ds->R > kGibbsMaxLagsFromSpikes)) {
        if (ds->x_dtype == PYGLM_B200_X_PLANES || ds->x_dtype == PYGLM_B200_X_NONE) {
            set_error("gibbs_begin: from-spikes mode needs F = 0 and R <= %d", kGibbsMaxLagsFromSpikes);
            return PYGLM_B200_EUNSUPPORTED;
        }
        spk = false;
    }
    cudaStream_t st = ds->stream;
    ds->gibbs_active = false;
    ds->g_spk = spk;
    TRY(stage_params(ds, bias, w, A, W, ds->g_bias, ds->g_w, ds->g_A, ds->g_W, st));
    const int ncols = n_hi - n_lo;
    TRY(ds->Inet.ensure((size_t)ncols * (ds->T ? ds->T : 1)));
    if (spk) {
        ds->g_nlo = n_lo; ds->g_ncols = ncols; ds->g_nlin = nlin;
        GibbsArgs g = gibbs_args(ds);
        TRY(launch_gibbs_inet_from_spikes(g, st));
    } else {
        if (!ds->xt_ready) {
            const size_t esz = ds->x_dtype == PYGLM_B200_X_F32 ? 4 : 8;
            const size_t NB = (size_t)ds->N * ds->B;
            TRY(ds->Xt.ensure(NB * (ds->T ? ds->T : 1) * esz));
            TRY(launch_transpose_X(ds->X.p, ds->T, (int64_t)NB, ds->ldx, ds->x_dtype, ds->Xt.p, st));
            ds->xt_ready = true;
        }
        TRY(ds->o_ll.ensure(ncols));
        // I_net[:, n] = I_imp @ (A[:,n] * W[:,n])  (glm.py:39) via the FP64 forward contraction
        TRY(ll_grad_dev_impl(ds, ds->g_bias.p, ds->g_w.p, ds->g_A.p, ds->g_W.p, nlin, n_lo, n_hi, PYGLM_B200_PATH_FP64,
                             ds->o_ll.p, nullptr, nullptr, ds->Inet.p, nullptr, st));
    }
    PYGLM_CUDA(cudaStreamSynchronize(st));
    ds->g_nlo = n_lo; ds->g_ncols = ncols; ds->g_nlin = nlin;
    ds->gibbs_active = true;
    return PYGLM_B200_OK;
}

static int gibbs_check_edges(const pyglm_b200_dataset* ds, int M, const int32_t* cols, const int32_t* pres)
{
    PYGLM_REQUIRE(M >= 0 && M <= 65535, "gibbs: batch size %d outside [0,65535]", M);
    PYGLM_REQUIRE(M == 0 || (cols && pres), "gibbs: null edge arrays");
    std::vector<char> seen(ds->g_ncols, 0);
    for (int m = 0; m < M; ++m) {
        PYGLM_REQUIRE(cols[m] >= ds->g_nlo && cols[m] < ds->g_nlo + ds->g_ncols, "gibbs: column %d not resident", cols[m]);
        PYGLM_REQUIRE(pres[m] >= 0 && pres[m] < ds->N, "gibbs: bad presynaptic index %d", pres[m]);
        PYGLM_REQUIRE(!seen[cols[m] - ds->g_nlo], "gibbs: column %d appears twice in one batch", cols[m]);
        seen[cols[m] - ds->g_nlo] = 1;
    }
    return PYGLM_B200_OK;
}

int pyglm_b200_gibbs_delta_ll(pyglm_b200_dataset* ds, int32_t M, const int32_t* cols, const int32_t* pres,
                              int32_t Q, const double* w_cand, double* out_ll)
{
    DS_GUARD(ds);
    if (!ds->gibbs_active) { set_error("gibbs_delta_ll before gibbs_begin"); return PYGLM_B200_ESTATE; }
    TRY(gibbs_check_edges(ds, M, cols, pres));
    PYGLM_REQUIRE(Q >= 1 && Q <= kMaxCand, "gibbs_delta_ll: Q=%d outside [1,%d]", Q, kMaxCand);
    if (M == 0) return PYGLM_B200_OK;
    PYGLM_REQUIRE(w_cand && out_ll, "gibbs_delta_ll: null argument");
    cudaStream_t st = ds->stream;
    if (ds->T == 0) { for (int i = 0; i < M * Q; ++i) out_ll[i] = 0.0; return PYGLM_B200_OK; }
    TRY(ds->g_cols.ensure(M)); TRY(ds->g_pres.ensure(M));
    TRY(ds->g_wcand.ensure((size_t)M * Q)); TRY(ds->g_out.ensure((size_t)M * Q));
    GibbsArgs g = gibbs_args(ds);
    TRY(ds->partial.ensure((size_t)M * g.nchunks * Q));
    g.partial = ds->partial.p;
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_cols.p, cols, M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_pres.p, pres, M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_wcand.p, w_cand, (size_t)M * Q * sizeof(double), cudaMemcpyHostToDevice, st));
    TRY(launch_gibbs_delta(g, M, ds->g_cols.p, ds->g_pres.p, Q, ds->g_wcand.p, ds->g_out.p, st));
    PYGLM_CUDA(cudaMemcpyAsync(out_ll, ds->g_out.p, (size_t)M * Q * sizeof(double), cudaMemcpyDeviceToHost, st));
    PYGLM_CUDA(cudaStreamSynchronize(st));
    return PYGLM_B200_OK;
}

int pyglm_b200_gibbs_delta_ll_dev(pyglm_b200_dataset* ds, int32_t M, const int32_t* d_cols, const int32_t* d_pres,
                                  int32_t Q, const double* d_w_cand, double* d_out_ll, void* stream)
{
    DS_GUARD(ds);
    if (!ds->gibbs_active) { set_error("gibbs_delta_ll before gibbs_begin"); return PYGLM_B200_ESTATE; }
    PYGLM_REQUIRE(M >= 0 && M <= 65535, "gibbs: batch size %d outside [0,65535]", M);
    PYGLM_REQUIRE(Q >= 1 && Q <= kMaxCand, "gibbs_delta_ll: Q=%d outside [1,%d]", Q, kMaxCand);
    if (M == 0) return PYGLM_B200_OK;
    PYGLM_REQUIRE(d_cols && d_pres && d_w_cand && d_out_ll, "gibbs_delta_ll_dev: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (ds->T == 0) { PYGLM_CUDA(cudaMemsetAsync(d_out_ll, 0, (size_t)M * Q * sizeof(double), st)); return PYGLM_B200_OK; }
    GibbsArgs g = gibbs_args(ds);
    TRY(ds->partial.ensure((size_t)M * g.nchunks * Q));
    g.partial = ds->partial.p;
    return launch_gibbs_delta(g, M, d_cols, d_pres, Q, d_w_cand, d_out_ll, st);
}

int pyglm_b200_gibbs_commit(pyglm_b200_dataset* ds, int32_t M, const int32_t* cols, const int32_t* pres,
                            const int8_t* a_new, const double* w_new)
{
    DS_GUARD(ds);
    if (!ds->gibbs_active) { set_error("gibbs_commit before gibbs_begin"); return PYGLM_B200_ESTATE; }
    TRY(gibbs_check_edges(ds, M, cols, pres));
    if (M == 0) return PYGLM_B200_OK;
    PYGLM_REQUIRE(a_new && w_new, "gibbs_commit: null argument");
    cudaStream_t st = ds->stream;
    TRY(ds->g_cols.ensure(M)); TRY(ds->g_pres.ensure(M));
    TRY(ds->g_anew.ensure(M)); TRY(ds->g_wnew.ensure(M));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_cols.p, cols, M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_pres.p, pres, M * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_anew.p, a_new, M, cudaMemcpyHostToDevice, st));
    PYGLM_CUDA(cudaMemcpyAsync(ds->g_wnew.p, w_new, M * sizeof(double), cudaMemcpyHostToDevice, st));
    GibbsArgs g = gibbs_args(ds);
    TRY(launch_gibbs_commit(g, M, ds->g_cols.p, ds->g_pres.p, ds->g_anew.p, ds->g_wnew.p, st));
    PYGLM_CUDA(cudaStreamSynchronize(st));
    return PYGLM_B200_OK;
}

int pyglm_b200_gibbs_get_state(const pyglm_b200_dataset* ds, int8_t* A, double* W)
{
    DS_GUARD(ds);
    if (!ds->gibbs_active) { set_error("gibbs_get_state before gibbs_begin"); return PYGLM_B200_ESTATE; }
    const size_t NN = (size_t)ds->N * ds->N;
    PYGLM_CUDA(cudaStreamSynchronize(ds->stream));
    if (A) PYGLM_CUDA(cudaMemcpy(A, ds->g_A.p, NN, cudaMemcpyDeviceToHost));
    if (W) PYGLM_CUDA(cudaMemcpy(W, ds->g_W.p, NN * sizeof(double), cudaMemcpyDeviceToHost));
    return PYGLM_B200_OK;
}

int pyglm_b200_gibbs_end(pyglm_b200_dataset* ds)
{
    DS_GUARD(ds);
    ds->gibbs_active = false;
    return PYGLM_B200_OK;
}

}  // extern "C"
