"""CPU tests of the oracle itself: against the committed golden vectors, against the scipy call the
reference makes, against torch.autograd (the definition T.grad implements), against finite
differences, and against the reference's own lam == f_nlin(X_sim) assertion."""
import os

import numpy as np
import pytest
import torch

from oracle import pyglm_oracle as orc
from tests.helpers import make_ibasis, make_problem, rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.mark.parametrize("name", ["standard_glm_n4.npz", "network_glm_n9_softplus.npz", "network_glm_n9_exp.npz"])
def test_oracle_reproduces_golden_ll_and_gradient(name):
    g = load(name)
    S = g['S'].astype(np.float64)
    fS = orc.convolve_with_basis(S, g['ibasis'])
    assert np.allclose(fS[::100], g['fS_rows'], rtol=0, atol=1e-12)
    ll, gb, gw = orc.population_ll_grad(fS, S, float(g['dt']), g['bias'], g['w'], g['A'], g['W'], int(g['nlin']))
    assert rel_err(ll, g['ll']) < 1e-12 and rel_err(gb, g['g_bias']) < 1e-10 and rel_err(gw, g['g_w']) < 1e-10
    # reference-shaped per-neuron path == whole-population GEMM path
    for n in range(S.shape[1]):
        l1, b1, w1 = orc.glm_ll_grad(fS, S, float(g['dt']), n, g['bias'][n], g['w'][n], g['A'], g['W'], int(g['nlin']))
        assert abs(l1 - g['ll'][n]) < 1e-9 * abs(g['ll'][n])
        assert np.allclose(w1, g['g_w'][n], rtol=1e-9, atol=1e-10)


def test_oracle_reproduces_golden_stimulus_fixture():
    g = load("stimulus_glm_n5.npz")
    S = g['S'].astype(np.float64)
    nT = S.shape[0]
    ib_s = orc.interpolate_stim_basis(orc.create_basis(dict(type='cosine', n_eye=0, n_cos=3, a=1.0 / 120, b=0.5,
                                                            orth=False, norm=True)), float(g['dt']), 0.3, True)
    assert np.allclose(ib_s, g['stim_ibasis'], rtol=0, atol=1e-15)
    istim, fstim = orc.filter_stimulus(g['stim'], float(g['dt_stim']), nT, float(g['dt']), ib_s)
    assert np.allclose(istim[::100], g['istim_rows'], rtol=0, atol=1e-13)
    assert np.allclose(fstim[::100], g['fstim_rows'], rtol=0, atol=1e-12)
    # the FFT filter of the real-valued stimulus == its defining causal sum
    direct = np.zeros_like(fstim)
    R, Bs = ib_s.shape
    for d in range(2):
        for b in range(Bs):
            direct[:, d * Bs + b] = np.convolve(istim[:, d], np.concatenate([[0.0], ib_s[:, b]]))[:nT]
    assert np.max(np.abs(direct - fstim)) < 1e-12 * np.max(np.abs(fstim))
    fS = orc.convolve_with_basis(S, g['ibasis'])
    ll, gb, gw, gs = orc.population_ll_grad(fS, S, float(g['dt']), g['bias'], g['w'], g['A'], g['W'], int(g['nlin']),
                                            fstim=fstim, w_stim=g['w_stim'])
    assert rel_err(ll, g['ll']) < 1e-12 and rel_err(gb, g['g_bias']) < 1e-10
    assert rel_err(gw, g['g_w']) < 1e-10 and rel_err(gs, g['g_w_stim']) < 1e-10
    # d ll / d w_stim by central differences
    eps = 1e-6
    for (n, f) in ((0, 0), (3, 5)):
        wp, wm = g['w_stim'].copy(), g['w_stim'].copy()
        wp[n, f] += eps
        wm[n, f] -= eps
        lp = orc.population_ll_grad(fS, S, float(g['dt']), g['bias'], g['w'], g['A'], g['W'], int(g['nlin']), fstim=fstim, w_stim=wp)[0][n]
        lm = orc.population_ll_grad(fS, S, float(g['dt']), g['bias'], g['w'], g['A'], g['W'], int(g['nlin']), fstim=fstim, w_stim=wm)[0][n]
        assert abs((lp - lm) / (2 * eps) - gs[n, f]) < 1e-5 * max(1.0, abs(gs[n, f]))


def test_golden_standard_glm_matches_simulator_activation():
    """The reference's own check (test/generate_synth_data.py:124-129): firing rate from the likelihood
    graph equals f_nlin of the activation the simulator accumulated in lag space."""
    g = load("standard_glm_n4.npz")
    S = g['S'].astype(np.float64)
    fS = orc.convolve_with_basis(S, g['ibasis'])
    x = orc.population_activation(fS, g['bias'], g['w'], g['A'], g['W'])
    assert np.allclose(orc.nlin(x[::100], orc.NLIN_SOFTPLUS), orc.nlin(g['Xsim_rows'], orc.NLIN_SOFTPLUS))
    assert np.max(np.abs(x[::100] - g['Xsim_rows'])) < 1e-9


def test_golden_gibbs_column_is_reproduced():
    g = load("gibbs_column_n5.npz")
    S = g['S']
    fS = orc.convolve_with_basis(S.astype(np.float64), g['ibasis'])
    A, W = g['A0'].copy(), g['W0'].copy()
    wn = g['wnorm']
    rec = orc.collapsed_column_sweep(fS, S, float(g['dt']), int(g['n_post']), g['bias'][int(g['n_post'])],
                                     g['w'][int(g['n_post'])], A, W, g['p_A'], orc.NLIN_SOFTPLUS, 0.0, 1.0, -0.2, 0.5,
                                     g['order'], g['uniforms'], lambda n_pre, a, mu, sig, ws, lL: mu + sig * wn[n_pre])
    assert np.array_equal(np.array([r['A'] for r in rec], dtype=np.int8), g['A_dec'])
    assert np.allclose([r['log_pr_A'] for r in rec], g['log_pr_A'], rtol=1e-12)
    assert np.array_equal(A, g['A_final']) and np.allclose(W, g['W_final'])


def test_filter_fft_form_equals_causal_sum():
    """utils/basis.py:220-234 (zero row + fftconvolve 'full'[:T]) == sum_{k>=1} basis[k-1] S[t-k]."""
    p = make_problem(2500, 5, 5)
    S = p['S'].astype(np.float64)
    a = orc.convolve_with_basis(S, p['ibasis'])
    b = orc.convolve_with_basis_direct(S, p['ibasis'])
    assert np.max(np.abs(a - b)) < 1e-12
    # brute force on a few bins, straight from the definition
    R = p['ibasis'].shape[0]
    for t in (0, 1, 7, 199, 200, 201, 2499):
        ref = np.zeros((5, 5))
        for k in range(1, R + 1):
            if t - k >= 0:
                ref += S[t - k][:, None] * p['ibasis'][k - 1][None, :]
        assert np.allclose(b[t], ref, atol=1e-12)
    assert np.all(b[0] == 0)          # causal: nothing reaches bin 0


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
def test_gradient_matches_autograd_of_the_reference_graph(nlin):
    """Write glm.py:33-52 / impulse.py:58 literally in torch float64 and differentiate it."""
    p = make_problem(1500, 6, 4, network=True, seed=3)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
    S = torch.from_numpy(p['S'].astype(np.float64))
    fS = torch.from_numpy(orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis']))
    A = torch.from_numpy(p['A'].astype(np.float64))
    W = torch.from_numpy(p['W'])
    fSn = fS.numpy()
    for n in (0, 3, 5):
        bias = torch.tensor([p['bias'][n]], requires_grad=True)
        w_ir = torch.tensor(p['w'][n].ravel(), requires_grad=True)
        w_ir3 = w_ir.reshape(1, 6, 4)
        I_imp = torch.sum(fS * w_ir3, dim=2)                           # impulse.py:58
        W_eff = A[:, n] * W[:, n]                                      # glm.py:33-35
        I_net = I_imp @ W_eff                                          # glm.py:39
        x = bias[0] + 0.0 + I_net
        lam = torch.exp(x) if nlin == orc.NLIN_EXP else torch.log(1.0 + torch.exp(x))   # nlin.py:25,43
        ll = torch.sum(-p['dt'] * lam + torch.log(lam) * S[:, n])      # glm.py:52
        ll.backward()
        l2, gb2, gw2 = orc.glm_ll_grad(fSn, p['S'], p['dt'], n, p['bias'][n], p['w'][n], p['A'], p['W'], nlin)
        assert abs(l2 - ll.item()) < 1e-10 * abs(ll.item())
        assert abs(gb2 - bias.grad.item()) < 1e-9 * max(1.0, abs(gb2))
        assert np.allclose(gw2.ravel(), w_ir.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_gradient_matches_finite_differences():
    p = make_problem(1200, 4, 5, network=True, seed=8)
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    n, eps = 2, 1e-6
    ll, gb, gw = orc.glm_ll_grad(fS, p['S'], p['dt'], n, p['bias'][n], p['w'][n], p['A'], p['W'], orc.NLIN_SOFTPLUS)
    f = lambda b, w: orc.glm_ll(fS, p['S'], p['dt'], n, b, w, p['A'], p['W'], orc.NLIN_SOFTPLUS)
    assert abs((f(p['bias'][n] + eps, p['w'][n]) - f(p['bias'][n] - eps, p['w'][n])) / (2 * eps) - gb) < 1e-5 * abs(gb)
    for (i, b) in ((0, 0), (1, 3), (3, 4)):
        wp, wm = p['w'][n].copy(), p['w'][n].copy()
        wp[i, b] += eps; wm[i, b] -= eps
        fd = (f(p['bias'][n], wp) - f(p['bias'][n], wm)) / (2 * eps)
        assert abs(fd - gw[i, b]) < 1e-5 * max(1.0, abs(gw[i, b]))


def test_dirichlet_chain_rule_matches_autograd():
    rng = np.random.default_rng(0)
    g = rng.gamma(1.0, 1.0, size=(5, 4)) * rng.choice([-1.0, 1.0], size=(5, 4))
    gbeta = rng.standard_normal((5, 4))
    gt = torch.tensor(g, requires_grad=True)
    beta = gt.abs() / gt.abs().sum(dim=1, keepdim=True)               # impulse.py:286-291
    (beta * torch.from_numpy(gbeta)).sum().backward()
    assert np.allclose(orc.dirichlet_chain_rule(g, gbeta), gt.grad.numpy(), rtol=1e-10)


def test_log_sum_exp_sample_rule():
    """First index whose cumulative probability reaches u (log_sum_exp.py:26-32)."""
    lnp = np.log([0.2, 0.5, 0.3]) + 1234.5
    assert orc.log_sum_exp_sample(lnp, 0.0) == 0
    assert orc.log_sum_exp_sample(lnp, 0.2 - 1e-12) == 0
    assert orc.log_sum_exp_sample(lnp, 0.2 + 1e-9) == 1
    assert orc.log_sum_exp_sample(lnp, 0.7 + 1e-9) == 2
    assert orc.log_sum_exp_sample([-np.inf, 0.0], 0.3) == 1
    with pytest.raises(Exception):
        orc.log_sum_exp_sample([-np.inf, -np.inf], 0.5)


def test_gauss_hermite_marginal_matches_dense_quadrature():
    """log G from 10 GH nodes (gibbs.py:1002-1022) integrates exp(log L(w)) N(w; mu, sigma^2) dw."""
    mu, sig = 0.3, 0.7
    f = lambda w: 0.8 * w - 1.5                               # E[exp(f(w))] has a closed form
    lp_noA, lp_A = orc.collapsed_edge_log_odds(f(orc.gh_candidates(mu, sig)), f(0.0), 0.25)
    exact = 0.8 * mu - 1.5 + 0.5 * (0.8 * sig) ** 2
    assert abs(lp_A - (np.log(0.25) + exact)) < 1e-9
    assert abs(lp_noA - (np.log(0.75) + f(0.0))) < 1e-12


def test_basis_shapes_and_normalisation():
    ib = make_ibasis(5)
    assert ib.shape == (200, 5)
    ibd = make_ibasis(5, kind="dirichlet")
    assert ibd.shape == (200, 5) and np.all(ibd >= 0)
    t = np.arange(0.0, 0.2, 0.001)
    assert np.allclose(np.trapezoid(ibd, t, axis=0), 1.0)               # impulse.py:373-374
    b = orc.create_basis(dict(type='cosine', n_eye=0, n_cos=5, a=1 / 120., b=0.5, orth=True, norm=False))
    assert np.allclose(b.T @ b, np.eye(5), atol=1e-12)                  # orthonormal columns (basis.py:97-98)
