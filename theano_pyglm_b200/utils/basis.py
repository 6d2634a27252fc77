"""Impulse-response bases (host side, NumPy) and the GPU spike-history filter entry point.

Mirrors the interface of pyglm/utils/basis.py: `create_basis(prms)` (:9) builds the 100-point
basis from the model dict; `convolve_with_basis(stim, basis)` (:201) is the filter, here executed
by CUDA kernel K1 instead of B FFT passes.  Basis construction is O(100 x B) and stays on the
host; the engine takes the interpolated basis as an input.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

N_PTS = 100     # resolution of the template basis (basis.py:69)


def _normalise_unit_integral(basis):
    if np.any(basis < 0):
        raise Exception("We can only normalize nonnegative impulse responses!")
    n_pts = basis.shape[0]
    return basis / basis.sum(axis=0, keepdims=True) * n_pts        # integral over [0,1] is 1 (basis.py:104)


def _finish(basis, prms):
    if prms.get('orth', False):
        basis = scipy.linalg.orth(basis)                            # basis.py:97-98
    if prms.get('norm', False):
        basis = _normalise_unit_integral(basis)
    return basis


def create_cosine_basis(prms):
    """Raised cosines with log-warped time axis; peaks spread over the first half (basis.py:56-106)."""
    n_eye, n_cos = prms['n_eye'], prms['n_cos']
    basis = np.zeros((N_PTS, n_eye + n_cos))
    basis[:n_eye, :n_eye] = np.eye(n_eye)
    warped = np.log(prms['a'] * np.arange(N_PTS) + prms['b'])
    peak_idx = np.floor(np.linspace(n_eye, N_PTS / 2.0, n_cos)).astype(int)
    centers = warped[peak_idx]
    width = centers / 2 if n_cos == 1 else (centers[-1] - centers[0]) / (n_cos - 1)
    for i, c in enumerate(np.atleast_1d(centers)):
        phase = np.clip((warped - c) * np.pi / width / 2.0, -np.pi, np.pi)
        basis[:, n_eye + i] = 0.5 * (np.cos(phase) + 1.0)
    return _finish(basis, prms)


def create_exp_basis(prms):
    """Decaying exponentials with log-spaced time constants (basis.py:108-143)."""
    n_eye, n_exp = prms['n_eye'], prms['n_exp']
    basis = np.zeros((N_PTS, n_eye + n_exp))
    basis[:n_eye, :n_eye] = np.eye(n_eye)
    taus = np.logspace(np.log10(1), np.log10(N_PTS / 3), n_exp)
    t = np.arange(N_PTS)
    for i, tau in enumerate(taus):
        basis[:, n_eye + i] = np.exp(-t / tau)
    return _finish(basis, prms)


def create_identity_basis(prms):
    return np.eye(prms['n_eye'])                                    # basis.py:188-199


def create_basis(prms):
    kind = prms['type'].lower()
    if kind == 'cosine':
        return create_cosine_basis(prms)
    if kind == 'exp':
        return create_exp_basis(prms)
    if kind in ('identity', 'eye'):
        return create_identity_basis(prms)
    raise Exception("Unrecognized basis type: %s" % kind)


def interpolate_basis(basis, dt, dt_max, norm, style="linear"):
    """Resample the template basis at the bin width of the data.

    style="linear"    LinearBasisImpulses.initialize_basis (impulse.py:92-112): unit interval,
                      R = dt_max/dt points by linspace, optional 1/dt_max scaling.
    style="dirichlet" DirichletImpulses.initialize_basis (impulse.py:359-376): sampled at
                      arange(0, dt_max, dt), optional trapezoid normalisation."""
    L, B = basis.shape
    if style == "linear":
        R = int(dt_max / dt)
        t_new, t_old = np.linspace(0, 1, R), np.linspace(0, 1, L)
    else:
        t_new, t_old = np.arange(0.0, dt_max, step=dt), np.linspace(0.0, dt_max, L)
    ib = np.column_stack([np.interp(t_new, t_old, basis[:, b]) for b in range(B)])
    if norm:
        ib = ib / dt_max if style == "linear" else ib / np.trapezoid(ib, t_new, axis=0)
    return ib


def make_standard_ibasis(B=5, dt=0.001, dt_max=0.2):
    """The interpolated basis of the standard_glm template (models/standard_glm.py:50-71)."""
    prms = dict(type='cosine', n_eye=0, n_cos=B, a=1.0 / 120, b=0.5, orth=True, norm=False)
    return interpolate_basis(create_basis(prms), dt, dt_max, prms['norm'], "linear")


def convolve_with_basis(stim, basis, device=0, x_dtype="f64"):
    """fS[t,d,b] = sum_{k=1..R} basis[k-1,b] * stim[t-k,d] on the GPU (kernel K1).

    Drop-in for pyglm.utils.basis.convolve_with_basis when `stim` holds spike counts (small
    non-negative integers): returns the (T, D, B) float64 array the reference stores as data['fS']."""
    from ..engine import Dataset
    ds = Dataset(stim, 1.0, basis, x_dtype=x_dtype, device=device)
    try:
        return ds.fS()
    finally:
        ds.close()
