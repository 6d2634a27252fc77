"""The oracle and the host mirror against outputs of THE REFERENCE ITSELF.

`tests/golden/ref_*.npz` were produced by `python -m oracle.ref_fixtures`, which imports and runs the
reference's own Python from /root/reference (oracle/ref_loader.py; Theano replaced by a torch float64
evaluator of the reference's graph definitions).  These tests need no GPU and no /root/reference.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyglm_oracle as orc
from theano_pyglm_b200.models import model_factory as mf
from theano_pyglm_b200.population import Population
from theano_pyglm_b200.utils import basis as hb
from theano_pyglm_b200.utils import packvec as hp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


BASES = {"std5": dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=True, norm=False),
         "std10": dict(type='cosine', n_eye=0, n_cos=10, a=1.0 / 120, b=0.5, orth=True, norm=False),
         "dir5": dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=False, norm=True),
         "stim3": dict(type='cosine', n_eye=0, n_cos=3, a=1.0 / 120, b=0.5, orth=False, norm=True)}


def _same_up_to_column_sign(a, b, tol):
    """scipy.linalg.orth fixes a basis only up to the sign of each column across LAPACK builds."""
    assert a.shape == b.shape
    for j in range(a.shape[1]):
        assert min(np.max(np.abs(a[:, j] - b[:, j])), np.max(np.abs(a[:, j] + b[:, j]))) <= tol


# -- a2 -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", sorted(BASES))
def test_create_basis_equals_reference(tag):
    ref = load("ref_numpy_units.npz")["basis_" + tag]
    _same_up_to_column_sign(orc.create_basis(BASES[tag]), ref, 1e-12)
    _same_up_to_column_sign(hb.create_basis(BASES[tag]), ref, 1e-12)


# -- a1 -------------------------------------------------------------------------------------------
def test_convolve_with_basis_equals_reference():
    f = load("ref_numpy_units.npz")
    S, ib = f["conv_S"].astype(np.float64), f["conv_ibasis"]
    assert np.array_equal(orc.convolve_with_basis(S, ib), f["conv_fS"])          # the same FFT call: bit-identical
    assert rel(orc.convolve_with_basis_direct(S, ib), f["conv_fS"]) < 1e-13      # the defining causal sum
    assert np.array_equal(orc.convolve_with_basis(f["conv_stim"], f["conv_stim_ibasis"]), f["conv_fstim"])
    assert rel(orc.convolve_with_basis_direct(f["conv_stim"], f["conv_stim_ibasis"]), f["conv_fstim"]) < 1e-13


# -- a7 -------------------------------------------------------------------------------------------
def test_nonlinearities_equal_reference():
    f = load("ref_numpy_units.npz")
    x = f["nlin_x"]
    assert rel(orc.nlin(x, orc.NLIN_EXP), f["nlin_exp"]) < 1e-15
    ref = f["nlin_explinear"]                        # np.log(1+np.exp(x)): overflows to inf at x = 710, 0 below -37
    ok = np.isfinite(ref) & (x > -30)
    assert ok.sum() > 140
    ours = orc.nlin(x, orc.NLIN_SOFTPLUS)
    # the literal form rounds 1+e^x to a multiple of 2^-52 before the log: agreement is to one ulp of 1.0 absolute
    assert np.all(np.abs(ours[ok] - ref[ok]) <= 2.3e-16 + 1e-14 * ref[ok])
    # where the reference's literal formula loses everything (1+e^x == 1) or overflows, the oracle keeps the limit
    assert np.all(np.isfinite(ours)) and ours[x == 700.0][0] == 700.0 and 0 < ours[x == -100.0][0] < 1e-40


# -- a14 ------------------------------------------------------------------------------------------
def test_log_sum_exp_sample_equals_reference():
    f = load("ref_numpy_units.npz")
    n_raise = 0
    for lnp, u, choice in zip(f["lse_lnp"], f["lse_u"], f["lse_choice"]):
        lnp = lnp[~np.isnan(lnp)]
        if choice < 0:
            n_raise += 1
            with pytest.raises(Exception):
                orc.log_sum_exp_sample(lnp, u)
        else:
            assert orc.log_sum_exp_sample(lnp, u) == choice
    assert len(f["lse_u"]) == 36 and n_raise < 6


def test_vectorised_decision_rule_equals_reference_decisions():
    """The lock-step sweep decides all columns' edges at once (`_decide_batch`: A = (u > p_noA)); fed the reference's
    own 11 `_glm_ll` values and uniforms it must reproduce the reference's A decisions edge for edge."""
    from theano_pyglm_b200.inference.gibbs import CollapsedGibbsNetworkColumnUpdate
    f = load("ref_sparse_weighted_n6.npz")
    upd = CollapsedGibbsNetworkColumnUpdate()
    for ci, n_post in enumerate(f["gibbs_cols"]):
        order = f["gibbs_order"][ci]
        ll = f["gibbs_glm_ll"][ci]
        u = f["gibbs_uniforms"][ci]
        pA = f["p_A"][order, n_post]
        with np.errstate(divide='ignore'):
            a = upd._decide_batch(ll, pA, u)
        assert np.array_equal(a, f["gibbs_A_after"][ci][order, n_post])


# -- a9: sorted-key packing -----------------------------------------------------------------------
def test_packdict_equals_reference():
    f = load("ref_numpy_units.npz")
    meta = json.loads(str(f["meta_json"]))
    nested = {'bias': {'bias': np.array([20.5])}, 'bkgd': {'w_stim': np.arange(6.0) * 0.1},
              'imp': {'w_ir': np.arange(10.0) - 4.0}, 'nlin': {}}
    vec, shapes = hp.packdict(nested)
    assert np.array_equal(vec, f["pack_vec"])
    assert json.loads(json.dumps(shapes)) == meta["pack_shapes"]
    assert np.array_equal(hp.unpackdict(vec * 2.0, shapes)['imp']['w_ir'], f["pack_back_w_ir"])
    imp12 = {'g_%d' % i: np.full(2, float(i)) for i in range(12)}
    assert np.array_equal(hp.packdict(imp12)[0], f["pack_vec_g12"])              # g_10, g_11 sort before g_2


# -- model dictionaries ---------------------------------------------------------------------------
def test_models_and_stabilize_sparsity_equal_reference():
    meta = json.loads(str(load("ref_numpy_units.npz")["meta_json"]))
    assert json.loads(json.dumps(mf.make_model('standard_glm', N=4, dt=0.001))) == meta["standard_glm_n4"]
    assert json.loads(json.dumps(mf.make_model('sparse_weighted_model', N=256, dt=0.001))) == meta["sparse_weighted_n256_before"]
    for N, rho in meta["stabilize_sparsity_rho"].items():
        m = mf.make_model('sparse_weighted_model', N=int(N), dt=0.001)
        mf.stabilize_sparsity(m)
        assert m['network']['graph']['rho'] == rho


# -- whole populations run by the reference ---------------------------------------------------------
POPS = ["ref_standard_glm_n4.npz", "ref_standard_glm_n3_exp.npz", "ref_sparse_weighted_n6.npz", "ref_stimulus_glm_n3.npz"]


def _state_from_fixture(f, model):
    """State dict x (reference layout, population.py:149-162) from the arrays of a population fixture."""
    N = model['N']
    glms = []
    for n in range(N):
        xn = {'n': n, 'bias': {'bias': np.array([f["bias"][n]])}, 'bkgd': {}, 'nlin': {}}
        if "w_ir" in f:
            xn['imp'] = {'w_ir': f["w_ir"][n].copy()}
        else:
            xn['imp'] = {'g_%d' % k: f["g"][n, k].copy() for k in range(N)}
        if "w_stim" in f:
            xn['bkgd'] = {'w_stim': f["w_stim"][n].copy()}
        glms.append(xn)
    net = {'graph': {}, 'weights': {}}
    if "A" in f:
        net = {'graph': {'A': f["A"].copy()}, 'weights': {'W': f["W"].reshape(-1).copy()}}
    return {'latent': {}, 'net': net, 'glms': glms}


def _oracle_blocks(f, model):
    """bias, w (N,N,B) [beta for Dirichlet], A, W, fS, fstim, w_stim in the oracle's conventions."""
    N = model['N']
    if "w_ir" in f:
        w = f["w_ir"].reshape(N, N, -1)
    else:
        w = np.stack([orc.dirichlet_beta(f["g"][n]) for n in range(N)])
    A = f["A"] if "A" in f else np.ones((N, N), dtype=np.int8)
    W = f["W"] if "W" in f else np.ones((N, N))
    S = f["S"].astype(np.float64)
    fS = orc.convolve_with_basis(S, f["ibasis"])
    fstim = w_stim = None
    if "w_stim" in f:
        _, fstim = orc.filter_stimulus(f["stim"], float(f["dt_stim"]), S.shape[0], float(f["dt"]), f["stim_ibasis"])
        w_stim = f["w_stim"]
    return f["bias"], w, A, W, S, fS, fstim, w_stim


@pytest.mark.parametrize("name", POPS)
def test_oracle_reproduces_reference_population(name):
    f = load(name)
    model = json.loads(str(f["model_json"]))
    N, dt = model['N'], float(f["dt"])
    nlin = orc.NLIN_EXP if str(f["nlin"]) == "exp" else orc.NLIN_SOFTPLUS
    # a2: interpolated basis (impulse.py:92-112 / :359-376)
    bp = model['impulse']['basis']
    interp = orc.interpolate_basis_dirichlet if model['impulse']['type'] == 'dirichlet' else orc.interpolate_basis_linear
    _same_up_to_column_sign(interp(orc.create_basis(bp), dt, model['impulse']['dt_max'], bp['norm']), f["ibasis"], 1e-12)
    bias, w, A, W, S, fS, fstim, w_stim = _oracle_blocks(f, model)
    # a1 through preprocess_data (impulse.py:114-130)
    assert np.array_equal(fS[::50], f["fS_rows"])
    if fstim is not None:
        assert rel(fstim[::50], f["fstim_rows"]) < 1e-12
    # a3-a8: glm.ll per neuron (glm.py:33-52)
    out = orc.population_ll_grad(fS, S, dt, bias, w, A, W, nlin, fstim=fstim, w_stim=w_stim)
    ll, gb, gw = out[0], out[1], out[2]
    assert rel(ll, f["ll"]) < 1e-11
    assert abs(float(np.sum(ll)) - float(f["total_ll"])) < 1e-9 * abs(float(f["total_ll"]))
    # the reference's own consistency assertion: lam(graph) == f_nlin(X_sim) (generate_synth_data.py:124-129)
    act = orc.population_activation(fS, bias, w, A, W) + (fstim @ w_stim.T if fstim is not None else 0.0)
    lam = orc.nlin(act, nlin)
    assert rel(lam[::50], f["lam_rows"]) < 1e-11
    if f["Xsim_rows"].size:
        assert rel(act[::50], f["Xsim_rows"]) < 1e-11
    # a9-a11: nlp / grad_nlp of coord_descent.py:40-80 in the reference's packed order
    for n in range(N):
        if model['impulse']['type'] == 'dirichlet':
            g = f["g"][n]
            lp = orc.bias_log_prior(bias[n], model['bias']['mu'], model['bias']['sigma']) + \
                orc.dirichlet_impulse_log_p(g, model['impulse']['alpha'])
            g_imp = orc.dirichlet_chain_rule(g, gw[n]) + ((model['impulse']['alpha'] - 1.0) / g - np.sign(g))
        else:
            pr = model['impulse']['prior']
            wn = w[n]
            lp = orc.bias_log_prior(bias[n], model['bias']['mu'], model['bias']['sigma']) + \
                orc.group_lasso_log_p(wn, pr['mu'], pr['sigma'], pr['lam'])
            g_imp = gw[n] + orc.group_lasso_log_p_grad(wn, pr['mu'], pr['sigma'], pr['lam'])
        parts = [np.array([gb[n] + orc.bias_log_prior_grad(bias[n], model['bias']['mu'], model['bias']['sigma'])])]
        if w_stim is not None:
            lp += orc.stim_log_prior(w_stim[n])
            parts.append(out[3][n] - w_stim[n] / 0.01 ** 2)
        parts.append(np.ravel(g_imp))
        assert abs(lp - f["log_prior_glm"][n]) < 1e-10 * max(1.0, abs(f["log_prior_glm"][n]))
        assert abs(-(lp + ll[n]) - f["nlp"][n]) < 1e-10 * abs(f["nlp"][n])
        assert rel(-np.concatenate(parts), f["grad_nlp"][n]) < 1e-9
    if "log_p_net" in f:
        lp_net = orc.erdos_renyi_log_p(A, f["p_A"]) + orc.gaussian_weight_log_p(W, 0.0, 1.0, -0.2, 0.5)
        assert abs(lp_net - float(f["log_p_net"])) < 1e-10 * abs(float(f["log_p_net"]))


@pytest.mark.parametrize("name", ["ref_standard_glm_n4.npz", "ref_standard_glm_n3_exp.npz", "ref_sparse_weighted_n6.npz"])
def test_host_mirror_reproduces_reference_prior_draws_and_simulation(name):
    """f3: with the same np.random seed, Population.sample() and .simulate() of the mirror consume the random stream in
    the reference's order and return the reference's parameters and spike trains bit for bit."""
    f = load(name)
    model = json.loads(str(f["model_json"]))
    N, dt = model['N'], float(f["dt"])
    popn = Population(model)
    np.random.seed(int(f["seed"]))
    x = popn.sample()
    assert np.array_equal(np.array([x['glms'][n]['bias']['bias'][0] for n in range(N)]), f["bias"])
    if "w_ir" in f:
        assert np.array_equal(np.stack([x['glms'][n]['imp']['w_ir'] for n in range(N)]), f["w_ir"])
    else:
        assert np.array_equal(np.stack([np.stack([x['glms'][n]['imp']['g_%d' % k] for k in range(N)]) for n in range(N)]), f["g"])
        assert np.array_equal(x['net']['graph']['A'], f["A"])
        assert np.array_equal(x['net']['weights']['W'].reshape(N, N), f["W"])
    S, X = popn.simulate(x, (0, float(f["T_sec"])), dt, None, None)
    assert np.array_equal(S.astype(np.uint8), f["S"])
    assert rel(X[::50], f["Xsim_rows"]) < 1e-12
    # packed parameter vector and prior of coord_descent.py in the reference's order
    for n in range(N):
        assert np.array_equal(popn.glm_param_vector(x['glms'][n]), f["x_vec"][n])
        assert abs(popn.glm.log_prior(x['glms'][n]) - f["log_prior_glm"][n]) < 1e-10 * max(1.0, abs(f["log_prior_glm"][n]))
    assert abs(popn.compute_log_prior(x) - float(f["total_log_prior"])) < 1e-10 * abs(float(f["total_log_prior"]))


def test_oracle_reproduces_reference_collapsed_gibbs_columns():
    """a12-a15: two `CollapsedGibbsNetworkColumnUpdate.update` calls run by the reference (shuffle order, 11 _glm_ll
    values per edge, uniforms, decisions; ARS replaced by the posterior-mode grid point in the fixture generator)."""
    f = load("ref_sparse_weighted_n6.npz")
    model = json.loads(str(f["model_json"]))
    N, dt = model['N'], float(f["dt"])
    bias, w, A0, W0, S, fS, _, _ = _oracle_blocks(f, model)
    A, W = A0.copy(), W0.copy()
    mu_w, sig_w, mu_ref, sig_ref = f["gibbs_mu_w"]
    for ci, n_post in enumerate(f["gibbs_cols"]):
        order = f["gibbs_order"][ci]
        u_all, z_all = list(f["gibbs_uniforms"][ci]), list(f["gibbs_randn"][ci])
        u_all = [u for u in u_all if not np.isnan(u)]
        z_all = [z for z in z_all if not np.isnan(z)]
        assert len(u_all) == N
        zi = iter(z_all)

        def w_draw(n_pre, a_new, mu, sig, W_nns, log_L):
            if a_new:                                   # the fixture's stand-in for ARS: argmax of log prior + log_L
                lp = -0.5 / sig ** 2 * (W_nns - mu) ** 2 + log_L
                ok = np.isfinite(lp) & (lp > -1e8)
                return W_nns[ok][np.argmax(lp[ok])]
            return mu + sig * next(zi)                  # gibbs.py:1063
        rec = orc.collapsed_column_sweep(fS, S, dt, int(n_post), bias[n_post], w[n_post], A, W, f["p_A"],
                                         orc.NLIN_SOFTPLUS, mu_w, sig_w, mu_ref, sig_ref, order, u_all, w_draw)
        ref_ll = f["gibbs_glm_ll"][ci]                  # (N edges, 11): 10 quadrature points then w = 0
        for i, r in enumerate(rec):
            # Where a candidate weight drives the rate to zero in a bin, the reference's literal log(1+e^x) underflows and
            # its ll is NaN (0 * -inf), which gibbs.py:1011-1012 turns into -inf; the stable form used by the oracle and
            # the engine keeps the (astronomically negative) finite value.  Both contribute nothing to log_G.
            fin = np.isfinite(ref_ll[i, :10])
            near = fin & (ref_ll[i, :10] > np.nanmax(ref_ll[i, :10]) - 300.0)     # the candidates that carry log_G
            assert near.sum() >= 1 and rel(r['log_L'][near], ref_ll[i, :10][near]) < 1e-11
            # further out the literal form has already lost digits (1 + e^x rounds for x < -30): 1e-4 there
            assert np.all(np.abs(r['log_L'][fin] - ref_ll[i, :10][fin]) <= 1e-4 * np.abs(ref_ll[i, :10][fin]))
            assert np.all(r['log_L'][~fin] < np.max(r['log_L']) - 100.0)
            assert abs(r['ll_noA'] - ref_ll[i, 10]) < 1e-11 * abs(ref_ll[i, 10])
        assert np.array_equal(A, f["gibbs_A_after"][ci])
        assert rel(W, f["gibbs_W_after"][ci]) < 1e-12
