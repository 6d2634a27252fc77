"""ctypes binding of the C ABI in include/pyglm_b200.h.

This is the only place Python touches the CUDA engine.  There is no CPU fallback: if the
shared library is missing or no GPU is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYGLM_B200_LIB") or os.path.join(_HERE, "lib", "libpyglm_b200.so")   # env: A/B builds

NLIN_EXP, NLIN_SOFTPLUS = 0, 1
X_F32, X_F64, X_PLANES, X_NONE = 0, 1, 2, 3
PATH_AUTO, PATH_FP64, PATH_TC = 0, 1, 2
_PATHS = {"auto": PATH_AUTO, "fp64": PATH_FP64, "tc": PATH_TC}
_NLINS = {"exp": NLIN_EXP, "explinear": NLIN_SOFTPLUS, "softplus": NLIN_SOFTPLUS}

EXPORTS = [
    "pyglm_b200_last_error", "pyglm_b200_abi_version",
    "pyglm_b200_dataset_create", "pyglm_b200_dataset_create_stim", "pyglm_b200_dataset_destroy",
    "pyglm_b200_dataset_num_stim", "pyglm_b200_filter_dense", "pyglm_b200_dataset_info",
    "pyglm_b200_dataset_get_fS", "pyglm_b200_dataset_device_X", "pyglm_b200_dataset_device_S",
    "pyglm_b200_dataset_refilter", "pyglm_b200_dataset_filter_bytes",
    "pyglm_b200_ll_grad", "pyglm_b200_ll_grad_dev", "pyglm_b200_resolve_path", "pyglm_b200_range_flags",
    "pyglm_b200_firing_rate",
    "pyglm_b200_gibbs_begin", "pyglm_b200_gibbs_delta_ll", "pyglm_b200_gibbs_delta_ll_dev", "pyglm_b200_gibbs_commit",
    "pyglm_b200_gibbs_get_state", "pyglm_b200_gibbs_end",
    "pyglm_b200_comm_create", "pyglm_b200_comm_export", "pyglm_b200_comm_connect", "pyglm_b200_allreduce_sum_dev",
    "pyglm_b200_comm_destroy", "pyglm_b200_comm_world", "pyglm_b200_ll_grad_allreduce_dev", "pyglm_b200_measure_fp64_peak",
]

_lib = None


class EngineError(RuntimeError):
    pass


def load_library():
    """dlopen the engine; fail loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            "pyglm_b200 CUDA engine not built: %s is missing. Run `python -m theano_pyglm_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.pyglm_b200_last_error.restype = C.c_char_p
    lib.pyglm_b200_last_error.argtypes = []
    lib.pyglm_b200_abi_version.restype = i32
    lib.pyglm_b200_dataset_create.argtypes = [p, i64, i32, i32, f64, p, i32, i32, i32, i32, C.POINTER(p)]
    lib.pyglm_b200_dataset_create_stim.argtypes = [p, i64, i32, i32, f64, p, i32, i32, p, i32, i32, i32, C.POINTER(p)]
    lib.pyglm_b200_dataset_num_stim.argtypes = [p]
    lib.pyglm_b200_dataset_num_stim.restype = i32
    lib.pyglm_b200_filter_dense.argtypes = [p, i64, i32, p, i32, i32, i32, p]
    lib.pyglm_b200_dataset_destroy.argtypes = [p]
    lib.pyglm_b200_dataset_info.argtypes = [p] + [p] * 7
    lib.pyglm_b200_dataset_get_fS.argtypes = [p, p]
    lib.pyglm_b200_dataset_device_X.argtypes = [p]
    lib.pyglm_b200_dataset_device_X.restype = p
    lib.pyglm_b200_dataset_device_S.argtypes = [p]
    lib.pyglm_b200_dataset_device_S.restype = p
    lib.pyglm_b200_dataset_refilter.argtypes = [p, p]
    lib.pyglm_b200_dataset_filter_bytes.argtypes = [p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.pyglm_b200_ll_grad.argtypes = [p, p, p, p, p, i32, i32, i32, i32, p, p, p]
    lib.pyglm_b200_ll_grad_dev.argtypes = [p, p, p, p, p, i32, i32, i32, i32, p, p, p, p]
    lib.pyglm_b200_resolve_path.argtypes = [p, i32]
    lib.pyglm_b200_range_flags.argtypes = [p, p]
    lib.pyglm_b200_firing_rate.argtypes = [p, p, p, p, p, i32, i32, i32, p]
    lib.pyglm_b200_gibbs_begin.argtypes = [p, p, p, p, p, i32, i32, i32]
    lib.pyglm_b200_gibbs_delta_ll.argtypes = [p, i32, p, p, i32, p, p]
    lib.pyglm_b200_gibbs_delta_ll_dev.argtypes = [p, i32, p, p, i32, p, p, p]
    lib.pyglm_b200_gibbs_commit.argtypes = [p, i32, p, p, p, p]
    lib.pyglm_b200_gibbs_get_state.argtypes = [p, p, p]
    lib.pyglm_b200_gibbs_end.argtypes = [p]
    lib.pyglm_b200_comm_create.argtypes = [i32, i32, i32, i64, C.POINTER(p)]
    lib.pyglm_b200_comm_export.argtypes = [p, p]
    lib.pyglm_b200_comm_connect.argtypes = [p, p]
    lib.pyglm_b200_allreduce_sum_dev.argtypes = [p, p, p, i64, p]
    lib.pyglm_b200_comm_destroy.argtypes = [p]
    lib.pyglm_b200_comm_world.argtypes = [p]
    lib.pyglm_b200_comm_world.restype = i32
    lib.pyglm_b200_ll_grad_allreduce_dev.argtypes = [p, p, p, p, p, p, i32, i32, p, p]
    lib.pyglm_b200_measure_fp64_peak.argtypes = [i32, p]
    for name in EXPORTS:      # every declared symbol must resolve
        getattr(lib, name)
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        msg = load_library().pyglm_b200_last_error().decode("utf-8", "replace")
        raise EngineError("pyglm_b200 error %d: %s" % (rc, msg))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        a = a.reshape(shape)
    return a


def spikes_to_u8(S):
    """Bin counts as uint8, bit-exact: data['S'] is float64 holding small non-negative integers
    (population.py:345-349 caps a bin at 10)."""
    S = np.asarray(S)
    if S.dtype == np.uint8:
        return np.ascontiguousarray(S)
    if S.size and (np.any(S < 0) or np.any(S > 255) or np.any(S != np.floor(S))):
        raise ValueError("spike counts must be integers in [0, 255]")
    return np.ascontiguousarray(S.astype(np.uint8))


def nlin_code(nlin):
    if isinstance(nlin, str):
        return _NLINS[nlin.lower()]
    return int(nlin)


def filter_dense(stim, ibasis, device=0):
    """convolve_with_basis for a real-valued signal (utils/basis.py:201-236; the stimulus projection of
    bkgd.py:145): stim (T, D), ibasis (R, B) -> (T, D, B) float64, computed on the GPU in FP64."""
    stim = _f64(stim)
    if stim.ndim == 1:
        stim = stim.reshape(-1, 1)
    ibasis = _f64(ibasis)
    T, D = stim.shape
    R, B = ibasis.shape
    out = np.empty((T, D * B))
    _check(load_library().pyglm_b200_filter_dense(_ptr(stim), T, D, _ptr(ibasis), R, B, int(device), _ptr(out)))
    return out.reshape(T, D, B)


def measure_fp64_peak(device=0):
    """FP64 FMA throughput of the GPU's CUDA cores, TFLOP/s (a DFMA loop; K4's roofline denominator)."""
    out = C.c_double()
    _check(load_library().pyglm_b200_measure_fp64_peak(int(device), C.addressof(out)))
    return float(out.value)


class Dataset:
    """One data sequence resident on one GPU: spikes + filtered spike train X (K1 output), optionally
    followed by F filtered-stimulus features `fstim` (T, F) (data['fstim'], bkgd.py:154)."""

    def __init__(self, S, dt, ibasis, halo=0, x_dtype="f32", device=0, fstim=None):
        lib = load_library()
        S = spikes_to_u8(S)
        if S.ndim != 2:
            raise ValueError("S must be (T, N)")
        ibasis = _f64(ibasis)
        if ibasis.ndim != 2:
            raise ValueError("ibasis must be (R, B)")
        self.T = int(S.shape[0]) - int(halo)
        self.N = int(S.shape[1])
        self.R, self.B = (int(v) for v in ibasis.shape)
        self.dt = float(dt)
        self.halo = int(halo)
        self.device = int(device)
        self.x_dtype = (X_F64 if x_dtype in ("f64", X_F64, np.float64) else
                        X_PLANES if x_dtype in ("planes", X_PLANES) else
                        X_NONE if x_dtype in ("none", "spikes", X_NONE) else X_F32)
        self.h2d_bytes = S.nbytes + ibasis.nbytes
        self.F = 0
        if fstim is not None and np.size(fstim):
            fstim = _f64(fstim)
            if fstim.ndim != 2 or fstim.shape[0] != self.T:
                raise ValueError("fstim must be (T, F) with T = %d bins" % self.T)
            self.F = int(fstim.shape[1])
            self.h2d_bytes += fstim.nbytes
        h = C.c_void_p()
        _check(lib.pyglm_b200_dataset_create_stim(_ptr(S), self.T, self.halo, self.N, self.dt, _ptr(ibasis),
                                                  self.R, self.B, _ptr(fstim) if self.F else None, self.F,
                                                  self.x_dtype, self.device, C.byref(h)))
        self._h = h
        ldx = C.c_int64()
        _check(lib.pyglm_b200_dataset_info(h, None, None, None, None, C.addressof(ldx), None, None))
        self.ldx = int(ldx.value)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            load_library().pyglm_b200_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- K1
    def fS(self):
        """data['fS'] (T, N, B) float64, copied back from the device (impulse.py:130)."""
        out = np.empty((self.T, self.N, self.B), dtype=np.float64)
        _check(load_library().pyglm_b200_dataset_get_fS(self._h, _ptr(out)))
        return out

    def refilter(self, stream=0):
        _check(load_library().pyglm_b200_dataset_refilter(self._h, C.c_void_p(stream)))

    def filter_bytes(self):
        """(bytes read, bytes written) by one K1 pass over this dataset."""
        r, w = C.c_int64(0), C.c_int64(0)
        _check(load_library().pyglm_b200_dataset_filter_bytes(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def device_X(self):
        return load_library().pyglm_b200_dataset_device_X(self._h)

    # -- K2
    def _params(self, bias, w, A, W, w_stim=None):
        N, NB = self.N, self.N * self.B
        bias = _f64(bias, (N,))
        w = _f64(w, (N, NB))
        if self.F:
            if w_stim is None:
                raise ValueError("this dataset carries %d stimulus features: pass w_stim (N, F)" % self.F)
            w = np.ascontiguousarray(np.concatenate([w, _f64(w_stim, (N, self.F))], axis=1))
        elif w_stim is not None and np.size(w_stim):
            raise ValueError("w_stim given but the dataset has no stimulus features")
        A = None if A is None else np.ascontiguousarray(A, dtype=np.int8).reshape(N, N)
        W = None if W is None else _f64(W, (N, N))
        return bias, w, A, W

    def ll_grad(self, bias, w, A=None, W=None, nlin="explinear", n_lo=0, n_hi=None, path="auto", grad=True,
                w_stim=None):
        """(ll, g_bias, g_w) for neurons [n_lo, n_hi); host buffers in and out (the e2e call).
        On a dataset with stimulus features, pass w_stim (N, F) and get (ll, g_bias, g_w, g_w_stim)."""
        n_hi = self.N if n_hi is None else n_hi
        nc = n_hi - n_lo
        bias, w, A, W = self._params(bias, w, A, W, w_stim)
        NB = self.N * self.B
        ll = np.empty(nc)
        gb = np.empty(nc) if grad else None
        gw = np.empty((nc, NB + self.F)) if grad else None
        _check(load_library().pyglm_b200_ll_grad(self._h, _ptr(bias), _ptr(w), _ptr(A), _ptr(W), nlin_code(nlin),
                                                 n_lo, n_hi, _PATHS.get(path, path), _ptr(ll), _ptr(gb), _ptr(gw)))
        if not grad:
            return ll
        if self.F:
            return ll, gb, np.ascontiguousarray(gw[:, :NB]), np.ascontiguousarray(gw[:, NB:])
        return ll, gb, gw

    def ll_grad_host_ptrs(self, bias, w, A, W, nlin, n_lo, n_hi, path, out_ll, out_gb, out_gw):
        """pyglm_b200_ll_grad on raw host addresses (integers; 0 = NULL), e.g. data_ptr() of pinned torch tensors:
        page-locked caller buffers are used in place, and the whole call replays as one CUDA graph."""
        vp = C.c_void_p
        _check(load_library().pyglm_b200_ll_grad(self._h, vp(bias), vp(w), vp(A) if A else None, vp(W) if W else None,
                                                 nlin_code(nlin), n_lo, n_hi, _PATHS.get(path, path), vp(out_ll),
                                                 vp(out_gb) if out_gb else None, vp(out_gw) if out_gw else None))

    def ll(self, bias, w, A=None, W=None, nlin="explinear", n_lo=0, n_hi=None, path="auto", w_stim=None):
        return self.ll_grad(bias, w, A, W, nlin, n_lo, n_hi, path, grad=False, w_stim=w_stim)

    def ll_grad_dev(self, d_bias, d_w, d_A, d_W, nlin, n_lo, n_hi, path, d_ll, d_gb, d_gw, stream):
        """Device-pointer variant: arguments are integer device addresses (tensor.data_ptr())."""
        vp = C.c_void_p
        _check(load_library().pyglm_b200_ll_grad_dev(self._h, vp(d_bias), vp(d_w), vp(d_A) if d_A else None,
                                                     vp(d_W) if d_W else None, nlin_code(nlin), n_lo, n_hi,
                                                     _PATHS.get(path, path), vp(d_ll), vp(d_gb) if d_gb else None,
                                                     vp(d_gw) if d_gw else None, vp(stream)))

    def ll_grad_allreduce_dev(self, comm, d_bias, d_w, d_A, d_W, nlin, path, d_out, stream):
        """Time-sharded evaluation: ll / gradients of all neurons on this rank's shard, summed over the ranks of `comm`
        (a PeerComm), into the device vector d_out = [ll | g_bias | g_w].  Device addresses as integers."""
        vp = C.c_void_p
        _check(load_library().pyglm_b200_ll_grad_allreduce_dev(self._h, comm._h if comm is not None else None, vp(d_bias), vp(d_w),
                                                               vp(d_A) if d_A else None, vp(d_W) if d_W else None,
                                                               nlin_code(nlin), _PATHS.get(path, path), vp(d_out), vp(stream)))

    def range_flags(self):
        """int32 (N,): 1 for every neuron whose activation left the FP32-safe range of the exp nonlinearity in the last
        tensor-core evaluation (the host call with path="auto" has already re-evaluated those in FP64)."""
        out = np.zeros(self.N, dtype=np.int32)
        _check(load_library().pyglm_b200_range_flags(self._h, _ptr(out)))
        return out

    def path_info(self, path="auto"):
        """What a call with `path` runs on this dataset (used by bench.py's roofline bookkeeping)."""
        use = load_library().pyglm_b200_resolve_path(self._h, _PATHS.get(path, path))
        if use == PATH_TC and self.N * self.B + self.F > 160:
            return dict(name="tcgen05-gemm-f16split", dtype="f16x2-split/f32-acc/f64-sum", x_passes=2,
                        launches_per_eval=5, bound="tensor", kernel="tc_gemm_fwd_kernel+tc_gemm_bwd_kernel")
        if use == PATH_TC:
            return dict(name="tcgen05-fused-f16split", dtype="f16x2-split/f32-acc/f64-sum", x_passes=1,
                        launches_per_eval=3, bound="hbm", kernel="tc_fused_kernel")
        if use == PATH_FP64:
            return dict(name="fp64-simt", dtype="f64", x_passes=2, launches_per_eval=5, bound="hbm",
                        kernel="simt_fwd_kernel+simt_bwd_kernel")
        raise EngineError("path %r unsupported for this dataset" % (path,))

    def firing_rate(self, bias, w, A=None, W=None, nlin="explinear", n_lo=0, n_hi=None, w_stim=None):
        n_hi = self.N if n_hi is None else n_hi
        bias, w, A, W = self._params(bias, w, A, W, w_stim)
        out = np.empty((self.T, n_hi - n_lo))
        _check(load_library().pyglm_b200_firing_rate(self._h, _ptr(bias), _ptr(w), _ptr(A), _ptr(W), nlin_code(nlin),
                                                     n_lo, n_hi, _ptr(out)))
        return out

    # -- K4
    def gibbs_begin(self, bias, w, A, W, nlin="explinear", n_lo=0, n_hi=None, w_stim=None):
        n_hi = self.N if n_hi is None else n_hi
        bias, w, A, W = self._params(bias, w, A, W, w_stim)
        _check(load_library().pyglm_b200_gibbs_begin(self._h, _ptr(bias), _ptr(w), _ptr(A), _ptr(W), nlin_code(nlin),
                                                     n_lo, n_hi))

    def gibbs_delta_ll(self, cols, pres, w_cand):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        pres = np.ascontiguousarray(pres, dtype=np.int32)
        w_cand = _f64(w_cand)
        if w_cand.ndim == 1:
            w_cand = w_cand.reshape(len(cols), -1)
        M, Q = w_cand.shape
        out = np.empty((M, Q))
        _check(load_library().pyglm_b200_gibbs_delta_ll(self._h, M, _ptr(cols), _ptr(pres), Q, _ptr(w_cand), _ptr(out)))
        return out

    def gibbs_delta_ll_dev(self, M, d_cols, d_pres, Q, d_w_cand, d_out, stream):
        """Device-pointer variant (integer addresses); enqueues on `stream`, no synchronisation."""
        vp = C.c_void_p
        _check(load_library().pyglm_b200_gibbs_delta_ll_dev(self._h, M, vp(d_cols), vp(d_pres), Q, vp(d_w_cand),
                                                            vp(d_out), vp(stream)))

    def gibbs_commit(self, cols, pres, a_new, w_new):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        pres = np.ascontiguousarray(pres, dtype=np.int32)
        a_new = np.ascontiguousarray(a_new, dtype=np.int8)
        w_new = _f64(w_new)
        _check(load_library().pyglm_b200_gibbs_commit(self._h, len(cols), _ptr(cols), _ptr(pres), _ptr(a_new), _ptr(w_new)))

    def gibbs_state(self):
        A = np.empty((self.N, self.N), dtype=np.int8)
        W = np.empty((self.N, self.N))
        _check(load_library().pyglm_b200_gibbs_get_state(self._h, _ptr(A), _ptr(W)))
        return A, W

    def gibbs_end(self):
        _check(load_library().pyglm_b200_gibbs_end(self._h))


class PeerComm:
    """One-shot sum all-reduce over NVLink peer memory between the ranks of a torch.distributed job
    (one process per GPU): the collective of time sharding.  `exchange` is a callable that takes this
    rank's 128 handle bytes and returns the list of every rank's bytes in rank order (see
    utils.parallel_util.make_peer_comm for the torch.distributed version)."""

    def __init__(self, rank, world, device, max_doubles, exchange):
        lib = load_library()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        _check(lib.pyglm_b200_comm_create(self.rank, self.world, self.device, int(max_doubles), C.byref(h)))
        self._h = h
        if self.world > 1:
            mine = (C.c_ubyte * 128)()
            _check(lib.pyglm_b200_comm_export(h, mine))
            parts = exchange(bytes(mine))
            if len(parts) != self.world or any(len(b) != 128 for b in parts):
                raise EngineError("peer handle exchange returned %d entries for world %d" % (len(parts), self.world))
            blob = (C.c_ubyte * (128 * self.world)).from_buffer_copy(b"".join(parts))
            _check(lib.pyglm_b200_comm_connect(h, blob))

    def allreduce_sum_dev(self, d_in, d_out, n, stream):
        """Device pointers (integers); asynchronous on `stream`; in place when d_in == d_out."""
        _check(load_library().pyglm_b200_allreduce_sum_dev(self._h, C.c_void_p(d_in), C.c_void_p(d_out), int(n),
                                                           C.c_void_p(stream)))

    def close(self):
        if getattr(self, "_h", None):
            load_library().pyglm_b200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
