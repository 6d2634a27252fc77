"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np

from oracle import pyglm_oracle as orc

STD_BASIS = dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=True, norm=False)
DIR_BASIS = dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=False, norm=True)


def make_ibasis(B=5, dt=0.001, dt_max=0.2, kind="standard"):
    if kind == "standard":      # standard_glm.py:63-71 + impulse.py:92-112
        prms = dict(STD_BASIS, n_cos=B)
        return orc.interpolate_basis_linear(orc.create_basis(prms), dt, dt_max, prms['norm'])
    prms = dict(DIR_BASIS, n_cos=B)   # sparse_weighted_model.py:51-63 + impulse.py:359-376
    return orc.interpolate_basis_dirichlet(orc.create_basis(prms), dt, dt_max, prms['norm'])


def make_problem(T, N, B=5, seed=1234, rate=0.02, network=False, R=None, dirichlet=False, multi=True):
    rng = np.random.default_rng(seed)
    ib = make_ibasis(B, kind="dirichlet" if dirichlet else "standard")
    if R is not None:
        ib = ib[:R].copy()
    S = (rng.random((T, N)) < rate).astype(np.uint8)
    if multi and T > 10:        # a few multi-spike bins: counts are integers, not just bits
        idx = rng.integers(0, T, size=max(1, T // 500))
        S[idx, rng.integers(0, N, size=idx.size)] = rng.integers(2, 5, size=idx.size)
    bias = 20.0 + 0.1 * rng.standard_normal(N)
    if dirichlet:
        g = rng.gamma(1.0, 1.0, size=(N, N, B))
        w = g / g.sum(axis=2, keepdims=True)
    else:
        w = 0.05 * rng.standard_normal((N, N, B))
    if network:
        A = (rng.random((N, N)) < 0.5).astype(np.int8)
        np.fill_diagonal(A, 1)
        W = rng.standard_normal((N, N))
        W[np.diag_indices(N)] = -0.2 + 0.5 * rng.standard_normal(N)
    else:
        A = np.ones((N, N), dtype=np.int8)
        W = np.ones((N, N))
    return dict(S=S, ibasis=ib, bias=bias, w=w, A=A, W=W, dt=0.001, T=T, N=N, B=B)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def grad_errors(got, ref, floor=1e-3):
    """(max-norm relative error, element-wise relative error over the entries with |ref| > floor * max|ref|)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    mx = max(np.max(np.abs(ref)), 1e-300)
    big = np.abs(ref) > floor * mx
    return float(np.max(np.abs(got - ref)) / mx), float(np.max(np.abs(got - ref)[big] / np.abs(ref)[big]))
