"""Generate the golden fixtures under tests/golden/ from the CPU oracle.

    python -m oracle.make_golden

The reference cannot run in this image (Python 2 + Theano 0.6), so the
fixtures are minted from oracle/pyglm_oracle.py on seeded inputs; the filter output in them
comes from the scipy.signal.fftconvolve call the reference itself makes (utils/basis.py:232).
The fixtures pin the oracle against silent drift and give the GPU tests fixed vectors.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyglm_oracle as orc  # noqa: E402
from tests.helpers import make_ibasis, make_problem  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def golden_standard_glm():
    """C1-shaped: standard_glm, N=4, softplus, complete graph, constant weights; spikes from the
    restated Population.simulate (population.py:233-389) rather than i.i.d. Bernoulli."""
    rng = np.random.default_rng(0)
    N, B, nT, dt = 4, 5, 3000, 0.001
    ib = make_ibasis(B)
    bias = 20.0 + 0.1 * rng.standard_normal(N)                       # standard_glm.py:16-21
    w = orc.sample_group_lasso(rng, N * N, B, 0.0, 10.0, 1.0).reshape(N, N, B) * 0.02
    A = np.ones((N, N), dtype=np.int8)
    W = np.ones((N, N))
    imps = np.einsum('npb,rb->pnr', w, ib)
    S, Xsim = orc.simulate(bias, imps, A, W, nT, dt, orc.NLIN_SOFTPLUS, rng)
    fS = orc.convolve_with_basis(S, ib)                              # the reference's FFT call
    ll, gb, gw = orc.population_ll_grad(fS, S, dt, bias, w, A, W, orc.NLIN_SOFTPLUS)
    lp_bias = np.array([orc.bias_log_prior(bias[n], 20.0, 0.1) for n in range(N)])
    lp_imp = np.array([orc.group_lasso_log_p(w[n], 0.0, 10.0, 1.0) for n in range(N)])
    np.savez_compressed(os.path.join(OUT, "standard_glm_n4.npz"),
                        S=S.astype(np.uint8), ibasis=ib, dt=dt, bias=bias, w=w, A=A, W=W,
                        Xsim_rows=Xsim[::100], fS_rows=fS[::100], ll=ll, g_bias=gb, g_w=gw,
                        lp_bias=lp_bias, lp_imp=lp_imp, nlin=orc.NLIN_SOFTPLUS)


def golden_network_glm():
    """C3-shaped: ER graph, Gaussian weights, Dirichlet impulses (beta-normalised), both nonlinearities."""
    for name, nlin, shift in (("softplus", orc.NLIN_SOFTPLUS, 0.0), ("exp", orc.NLIN_EXP, -17.0)):
        p = make_problem(2500, 9, 5, seed=42, network=True, dirichlet=True)
        p['bias'] = p['bias'] + shift
        fS = orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis'])
        ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin)
        np.savez_compressed(os.path.join(OUT, "network_glm_n9_%s.npz" % name),
                            S=p['S'], ibasis=p['ibasis'], dt=p['dt'], bias=p['bias'], w=p['w'], A=p['A'], W=p['W'],
                            fS_rows=fS[::100], ll=ll, g_bias=gb, g_w=gw, nlin=nlin)


def golden_gibbs_column():
    """One collapsed-Gibbs column sweep (gibbs.py:1229-1250) with injected shuffle/uniforms/weights."""
    T, N, B = 2000, 5, 5
    p = make_problem(T, N, B, seed=7, network=True, dirichlet=True)
    rng = np.random.default_rng(11)
    fS = orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis'])
    p_A = np.full((N, N), 0.3)
    np.fill_diagonal(p_A, 1.0 - 1e-3)
    n_post = 2
    order = rng.permutation(N)
    unif = rng.random(N)
    wn = rng.standard_normal(N)
    A, W = p['A'].copy(), p['W'].copy()
    rec = orc.collapsed_column_sweep(fS, p['S'], p['dt'], n_post, p['bias'][n_post], p['w'][n_post], A, W, p_A,
                                     orc.NLIN_SOFTPLUS, 0.0, 1.0, -0.2, 0.5, order, unif,
                                     lambda n_pre, a, mu, sig, ws, lL: mu + sig * wn[n_pre])
    np.savez_compressed(os.path.join(OUT, "gibbs_column_n5.npz"),
                        S=p['S'], ibasis=p['ibasis'], dt=p['dt'], bias=p['bias'], w=p['w'], A0=p['A'], W0=p['W'],
                        p_A=p_A, n_post=n_post, order=order, uniforms=unif, wnorm=wn,
                        log_L=np.array([r['log_L'] for r in rec]), ll_noA=np.array([r['ll_noA'] for r in rec]),
                        log_pr_A=np.array([r['log_pr_A'] for r in rec]),
                        log_pr_noA=np.array([r['log_pr_noA'] for r in rec]),
                        A_dec=np.array([r['A'] for r in rec], dtype=np.int8), A_final=A, W_final=W)


def golden_stimulus_glm():
    """standard_glm with a BasisStimulus background (bkgd.py:45-172): a 2-D stimulus sampled at 10 ms,
    interpolated to the bins, projected on a 3-function basis, weights w_stim (N, 6)."""
    rng = np.random.default_rng(5)
    N, B, nT, dt, dt_stim = 5, 5, 2500, 0.001, 0.01
    p = make_problem(nT, N, B, seed=77, network=True)
    stim = np.cumsum(rng.standard_normal((int(np.ceil(nT * dt / dt_stim)) + 1, 2)), axis=0) * 0.1
    prms = dict(type='cosine', n_eye=0, n_cos=3, a=1.0 / 120, b=0.5, orth=False, norm=True)
    ib_s = orc.interpolate_stim_basis(orc.create_basis(prms), dt, 0.3, True)
    istim, fstim = orc.filter_stimulus(stim, dt_stim, nT, dt, ib_s)
    w_stim = 0.05 * rng.standard_normal((N, fstim.shape[1]))
    fS = orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis'])
    ll, gb, gw, gs = orc.population_ll_grad(fS, p['S'], dt, p['bias'], p['w'], p['A'], p['W'], orc.NLIN_SOFTPLUS,
                                            fstim=fstim, w_stim=w_stim)
    np.savez_compressed(os.path.join(OUT, "stimulus_glm_n5.npz"),
                        S=p['S'], ibasis=p['ibasis'], dt=dt, bias=p['bias'], w=p['w'], A=p['A'], W=p['W'],
                        stim=stim, dt_stim=dt_stim, stim_ibasis=ib_s, istim_rows=istim[::100], fstim_rows=fstim[::100],
                        w_stim=w_stim, ll=ll, g_bias=gb, g_w=gw, g_w_stim=gs, nlin=orc.NLIN_SOFTPLUS,
                        lp_stim=np.array([orc.stim_log_prior(w_stim[n]) for n in range(N)]))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-new" not in sys.argv:      # the first four fixtures are frozen; regenerate them only on purpose
        golden_standard_glm()
        golden_network_glm()
        golden_gibbs_column()
    golden_stimulus_glm()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
