"""GPU tests of the callers on either side of the hot path: MAP coordinate descent (synth_map) and the
collapsed Gibbs sweep (synth_mcmc), through the Population interface."""
import copy

import numpy as np
import pytest
import scipy.optimize as opt

from oracle import pyglm_oracle as orc
from tests.helpers import rel_err
from theano_pyglm_b200.models.model_factory import make_model, stabilize_sparsity
from theano_pyglm_b200.population import Population

pytestmark = pytest.mark.gpu


def synth_standard_glm(N=4, nT=20000, seed=0):
    """Restated test/generate_synth_data.py for standard_glm: prior draw, simulate, package."""
    model = make_model('standard_glm', N=N, dt=0.001)
    popn = Population(model)
    np.random.seed(seed)
    x_true = popn.sample()
    for n in range(N):                                   # group-lasso draws are O(10); keep the network stable
        x_true['glms'][n]['imp']['w_ir'] *= 0.02
    rng = np.random.default_rng(seed)
    bias, w, A, W = popn.glm.engine_params(x_true)
    ib = popn.glm.imp_model.ibasis
    imps = np.einsum('npb,rb->pnr', w.reshape(N, N, -1), ib)
    S, X = orc.simulate(bias, imps, np.ones((N, N), np.int8), np.ones((N, N)), nT, 0.001, orc.NLIN_SOFTPLUS, rng)
    data = {'S': S, 'X': X, 'N': N, 'dt': 0.001, 'T': nT * 0.001, 'stim': None, 'dt_stim': 0.1, 'vars': x_true}
    return model, popn, data, x_true


def test_population_ll_matches_oracle_and_reference_assertion():
    model, popn, data, x_true = synth_standard_glm()
    popn.add_data(data)
    N = model['N']
    bias, w, A, W = popn.glm.engine_params(x_true)
    fS = orc.convolve_with_basis(data['S'], popn.glm.imp_model.ibasis)
    ll, gb, gw = orc.population_ll_grad(fS, data['S'], 0.001, bias, w.reshape(N, N, -1), np.ones((N, N), np.int8),
                                        np.ones((N, N)), orc.NLIN_SOFTPLUS)
    assert abs(popn.compute_ll(x_true) - ll.sum()) < 1e-6 * abs(ll.sum())
    lp = popn.compute_log_p(x_true)
    assert abs(lp - (popn.compute_log_prior(x_true) + ll.sum())) < 1e-6 * abs(lp)
    state = popn.eval_state(x_true)                      # generate_synth_data.py:124-129
    for n in range(N):
        assert np.allclose(state['glms'][n]['lam'], orc.nlin(data['X'][:, n], orc.NLIN_SOFTPLUS))
    # per-neuron log posterior + gradient vector in sorted-key order (bias, w_ir)
    lp1, g1 = popn.glm_log_p_grad(x_true, 1)
    ref_lp = ll[1] + popn.glm.log_prior(x_true['glms'][1])
    gp = popn.glm.grad_log_prior(x_true['glms'][1])
    ref_g = np.concatenate([[gb[1] + gp['bias']['bias'][0]], gw[1].ravel() + gp['imp']['w_ir']])
    assert abs(lp1 - ref_lp) < 1e-6 * abs(ref_lp) and rel_err(g1, ref_g) < 1e-5
    lps, gs = popn.glms_log_p_grad(x_true)
    assert abs(lps[1] - lp1) < 1e-9 * abs(lp1) and np.allclose(gs[1], g1, rtol=1e-9, atol=1e-9)


def test_map_coordinate_descent_batched_and_per_neuron_agree_with_cpu_optimum():
    from theano_pyglm_b200.inference.coord_descent import coord_descent, fit_glm
    model, popn, data, x_true = synth_standard_glm(N=3, nT=15000, seed=2)
    popn.add_data(data)
    N = 3
    np.random.seed(11)
    x0 = popn.sample()
    for n in range(N):
        x0['glms'][n]['imp']['w_ir'] *= 0.01
    lp0 = popn.compute_log_p(x0)
    xa = coord_descent(popn, x0=copy.deepcopy(x0), maxiter=1, batched=True)
    lpa = popn.compute_log_p(xa)
    assert lpa > lp0
    xb = copy.deepcopy(x0)
    for n in range(N):
        fit_glm(popn, xb, n)
    lpb = popn.compute_log_p(xb)
    # same optimum from the CPU oracle objective with scipy BFGS (the reference's optimiser)
    fS = orc.convolve_with_basis(data['S'], popn.glm.imp_model.ibasis)
    A, W = np.ones((N, N), np.int8), np.ones((N, N))
    lpc = popn.network.log_p(x0['net'])
    for n in range(N):
        def nll(v):
            ll, gb, gw = orc.glm_ll_grad(fS, data['S'], 0.001, n, v[0], v[1:].reshape(N, -1), A, W, orc.NLIN_SOFTPLUS)
            lp = ll + orc.bias_log_prior(v[0], 20, 0.1) + orc.group_lasso_log_p(v[1:].reshape(N, -1), 0.0, 10.0, 1.0)
            g = np.concatenate([[gb + orc.bias_log_prior_grad(v[0], 20, 0.1)],
                                (gw + orc.group_lasso_log_p_grad(v[1:].reshape(N, -1), 0.0, 10.0, 1.0)).ravel()])
            return -lp, -g
        v0 = popn.glm_param_vector(x0['glms'][n])
        res = opt.minimize(nll, v0, jac=True, method="bfgs", options={'maxiter': 225})
        lpc += -res.fun
    assert abs(lpa - lpc) < 2e-6 * abs(lpc)
    assert abs(lpb - lpc) < 2e-6 * abs(lpc)
    # and the fit moved toward the truth: higher likelihood than the generating parameters' prior draw
    assert lpa >= popn.compute_log_p(x_true) - 1e-3 * abs(lpa)


def synth_network_glm(N=5, nT=6000, seed=3):
    model = make_model('sparse_weighted_model', N=N, dt=0.001)
    stabilize_sparsity(model)
    popn = Population(model)
    np.random.seed(seed)
    x = popn.sample()
    rng = np.random.default_rng(seed)
    S = (rng.random((nT, N)) < 0.02).astype(np.float64)
    data = {'S': S, 'N': N, 'dt': 0.001, 'T': nT * 0.001, 'stim': None, 'dt_stim': 0.1}
    popn.add_data(data)
    return model, popn, data, x


def test_collapsed_column_update_reproduces_reference_decisions_for_the_same_random_stream():
    """`update(x, n)` consumes np.random where the reference does (shuffle, one rand per edge, one randn
    when W comes from the prior).  Replaying the same stream through the CPU restatement of
    gibbs.py:1229-1250 must give identical A decisions and the same W."""
    from theano_pyglm_b200.inference.gibbs import CollapsedGibbsNetworkColumnUpdate
    model, popn, data, x = synth_network_glm()
    N = model['N']
    upd = CollapsedGibbsNetworkColumnUpdate()
    upd.preprocess(popn)
    upd.sample_w_with_ars = False                         # ARS comes from the un-vendored hips: parity unpinned
    x_ref = copy.deepcopy(x)
    bias, w, A0, W0 = popn.glm.engine_params(x_ref)
    fS = orc.convolve_with_basis(data['S'], popn.glm.imp_model.ibasis)
    p_A = popn.network.graph.pA.get_value()
    A_ref, W_ref = A0.copy(), W0.copy()
    upd.begin(x)
    for n_post in (0, 3):
        np.random.seed(100 + n_post)
        upd.update(x, n_post)
        np.random.seed(100 + n_post)                      # replay
        order = np.arange(N)
        np.random.shuffle(order)
        us, zs = np.zeros(N), np.zeros(N)
        for i in range(N):
            us[i] = np.random.rand()
            zs[i] = np.random.randn()
        zmap = {int(p): zs[i] for i, p in enumerate(order)}
        orc.collapsed_column_sweep(fS, data['S'], 0.001, n_post, bias[n_post], w[n_post].reshape(N, -1), A_ref, W_ref,
                                   p_A, orc.NLIN_SOFTPLUS, 0.0, 1.0, -0.2, 0.5, order, us,
                                   lambda n_pre, a, mu, sig, ws, lL: mu + sig * zmap[int(n_pre)])
        assert np.array_equal(x['net']['graph']['A'], A_ref)
        assert np.allclose(x['net']['weights']['W'].reshape(N, N), W_ref, rtol=0, atol=1e-12)
    A_dev, W_dev = popn._handle().gibbs_state()
    assert np.array_equal(A_dev, A_ref) and np.allclose(W_dev, W_ref, atol=1e-12)
    upd.end()


def test_collapsed_column_update_conditions_on_every_data_sequence():
    """gibbs.py:931-935 sums `_glm_ll` over population.data_sequences.  With two sequences of different lengths the
    mirror's candidate log-likelihoods must be the sum of the oracle's per-sequence values, its decisions those of the
    summed log-likelihoods, and both resident states must receive every commit."""
    from theano_pyglm_b200.inference.gibbs import CollapsedGibbsNetworkColumnUpdate
    model, popn, data, x = synth_network_glm(N=5, nT=3000, seed=5)
    N = model['N']
    rng = np.random.default_rng(17)
    data2 = {'S': (rng.random((1700, N)) < 0.03).astype(np.float64), 'N': N, 'dt': 0.001, 'T': 1.7, 'stim': None, 'dt_stim': 0.1}
    popn.add_data(data2, set_as_current_data=False)        # `current` stays the first one: must not matter
    upd = CollapsedGibbsNetworkColumnUpdate()
    upd.preprocess(popn)
    upd.sample_w_with_ars = False
    bias, w, A0, W0 = popn.glm.engine_params(copy.deepcopy(x))
    ib = popn.glm.imp_model.ibasis
    seqs = [(orc.convolve_with_basis(d['S'], ib), d['S']) for d in (data, data2)]
    p_A = popn.network.graph.pA.get_value()
    ds = upd.begin(x)
    n_post, n_pre = 3, 1
    cand = upd._candidates(n_pre, n_post)
    got = ds.gibbs_delta_ll([n_post], [n_pre], cand[None, :])[0]
    ref = np.zeros(11)
    for fS, S in seqs:
        I_imp = orc.impulse_current(fS, w[n_post].reshape(N, -1))
        A1 = A0.copy(); A1[n_pre, n_post] = 0
        I_other = I_imp @ orc.effective_weights(A1, W0, n_post)
        ref += [orc.gibbs_glm_ll(bias[n_post], 0.0, I_other, I_imp[:, n_pre], wq, S[:, n_post], 0.001, orc.NLIN_SOFTPLUS)
                for wq in cand]
    assert rel_err(got, ref) < 1e-10
    # a whole column, replayed on the oracle with summed log-likelihoods
    np.random.seed(42)
    upd.update(x, n_post)
    np.random.seed(42)
    order = np.arange(N)
    np.random.shuffle(order)
    A_ref, W_ref = A0.copy(), W0.copy()
    for pre in order:
        u, z = np.random.rand(), np.random.randn()
        mu, sig = (-0.2, 0.5) if pre == n_post else (0.0, 1.0)
        W_nns = orc.gh_candidates(mu, sig)
        ll = np.zeros(11)
        for fS, S in seqs:
            I_imp = orc.impulse_current(fS, w[n_post].reshape(N, -1))
            A1 = A_ref.copy(); A1[pre, n_post] = 0
            I_other = I_imp @ orc.effective_weights(A1, W_ref, n_post)
            ll += [orc.gibbs_glm_ll(bias[n_post], 0.0, I_other, I_imp[:, pre], wq, S[:, n_post], 0.001, orc.NLIN_SOFTPLUS)
                   for wq in list(W_nns) + [0.0]]
        lp_no, lp_a = orc.collapsed_edge_log_odds(ll[:10], ll[10], p_A[pre, n_post])
        A_ref[pre, n_post] = orc.log_sum_exp_sample([lp_no, lp_a], u)
        W_ref[pre, n_post] = mu + sig * z
    assert np.array_equal(x['net']['graph']['A'], A_ref)
    assert np.allclose(x['net']['weights']['W'].reshape(N, N), W_ref, rtol=0, atol=1e-12)
    for d in (data, data2):
        A_dev, W_dev = popn._handle(d).gibbs_state()
        assert np.array_equal(A_dev, A_ref) and np.allclose(W_dev, W_ref, atol=1e-12)
    upd.end()


def test_gibbs_sample_runs_and_keeps_a_valid_state():
    from theano_pyglm_b200.inference.gibbs import gibbs_sample
    model, popn, data, x = synth_network_glm(N=4, nT=4000, seed=5)
    np.random.seed(7)
    for batched in (True, False):
        smpls = gibbs_sample(popn, N_samples=2, x0=copy.deepcopy(x), batched=batched)
        assert len(smpls) == 3
        last = smpls[-1]
        A = last['net']['graph']['A']
        assert A.dtype == np.int8 and set(np.unique(A)) <= {0, 1} and np.all(np.diag(A) == 1)
        assert np.isfinite(popn.compute_log_p(last))
        assert any(not np.array_equal(smpls[0]['net']['weights']['W'], s['net']['weights']['W']) for s in smpls[1:])
        # engine state and host state stayed in sync through the sweep
        assert np.all(np.isfinite(last['net']['weights']['W']))


def test_gibbs_sample_initialised_from_the_map_estimate():
    """gibbs.py:2486-2507: fit a standard GLM by coordinate descent on the same data, project it onto the
    weighted Dirichlet model (convert_model) and start the chain there."""
    from theano_pyglm_b200.inference.gibbs import gibbs_sample, initial_state
    model, popn, data, x = synth_network_glm(N=4, nT=6000, seed=9)
    np.random.seed(3)
    x0 = initial_state(popn, init_from_mle=True)
    N = 4
    assert x0['net']['graph']['A'].shape == (N, N) and x0['net']['graph']['A'].dtype == np.int8
    for n in range(N):
        for k in range(N):
            g = x0['glms'][n]['imp']['g_%d' % k]
            assert g.shape == (popn.glm.imp_model.B,) and np.all(g > 0)
    np.random.seed(3)
    lp_prior_draw = popn.compute_log_p(popn.sample())
    assert np.isfinite(popn.compute_log_p(x0))
    assert popn.compute_ll(x0) > popn.compute_ll(x) - abs(popn.compute_ll(x))      # a sane starting point, not a blow-up
    smpls = gibbs_sample(popn, N_samples=1, init_from_mle=True)
    assert len(smpls) == 2 and np.isfinite(popn.compute_log_p(smpls[-1])) and np.isfinite(lp_prior_draw)


def test_basis_stimulus_population_end_to_end():
    """A standard GLM with a BasisStimulus background (bkgd.py:45-172): simulate with a stimulus, then the
    reference's own assertion (lam from the likelihood graph == f_nlin of the simulated activation,
    generate_synth_data.py:124-129), the oracle's ll, and the gradient wrt (bias, w_stim, w_ir)."""
    N, T_stop = 3, 4.0
    model = make_model('standard_glm', N=N, dt=0.001)
    model['bkgd'] = {'type': 'basis', 'D_stim': 2, 'dt_max': 0.3, 'dt_stim': 0.1,
                     'basis': dict(type='cosine', n_eye=0, n_cos=3, a=1 / 120., b=0.5, orth=False, norm=True)}
    popn = Population(model)
    np.random.seed(4)
    x = popn.sample()
    for n in range(N):
        x['glms'][n]['imp']['w_ir'] *= 0.02
        x['glms'][n]['bkgd']['w_stim'] = 2.0 * np.random.randn(6)      # a stimulus drive worth a few Hz
    stim = np.random.randn(int(T_stop / 0.1), 2)
    S, X = popn.simulate(x, (0, T_stop), 0.001, stim, 0.1)
    data = {'S': S, 'X': X, 'N': N, 'dt': 0.001, 'T': T_stop, 'stim': stim, 'dt_stim': 0.1}
    popn.add_data(data)
    assert data['fstim'].shape == (4000, 6)
    # stimulus filtering == the oracle's restatement of bkgd.py:122-154
    ib_s = orc.interpolate_stim_basis(orc.create_basis(model['bkgd']['basis']), 0.001, 0.3, True)
    _, fstim = orc.filter_stimulus(stim, 0.1, 4000, 0.001, ib_s)
    assert np.max(np.abs(data['fstim'] - fstim)) < 1e-12 * max(1.0, np.max(np.abs(fstim)))
    state = popn.eval_state(x)
    for n in range(N):
        assert np.allclose(state['glms'][n]['lam'], orc.nlin(X[:, n], orc.NLIN_SOFTPLUS), rtol=1e-6, atol=1e-8)
    bias, w, A, W = popn.glm.engine_params(x)
    ws = popn.glm.stim_weights(x)
    fS = orc.convolve_with_basis(S, popn.glm.imp_model.ibasis)
    ll, gb, gw, gs = orc.population_ll_grad(fS, S, 0.001, bias, w.reshape(N, N, -1), np.ones((N, N), np.int8),
                                            np.ones((N, N)), orc.NLIN_SOFTPLUS, fstim=fstim, w_stim=ws)
    assert abs(popn.compute_ll(x) - ll.sum()) < 1e-6 * abs(ll.sum())
    lp1, g1 = popn.glm_log_p_grad(x, 1)
    gp = popn.glm.grad_log_prior(x['glms'][1])
    ref_g = np.concatenate([[gb[1] + gp['bias']['bias'][0]], gs[1] + gp['bkgd']['w_stim'],
                            gw[1].ravel() + gp['imp']['w_ir']])
    assert abs(lp1 - (ll[1] + popn.glm.log_prior(x['glms'][1]))) < 1e-6 * abs(lp1)
    assert rel_err(g1, ref_g) < 1e-5
    lps, gsb = popn.glms_log_p_grad(x)
    assert np.allclose(gsb[1], g1, rtol=1e-9, atol=1e-9)
    # MAP refit moves the stimulus weights towards the truth from zero
    from theano_pyglm_b200.inference.coord_descent import coord_descent
    x0 = copy.deepcopy(x)
    for n in range(N):
        x0['glms'][n]['bkgd']['w_stim'] = np.zeros(6)
    lp0 = popn.compute_log_p(x0)
    x_fit = coord_descent(popn, x0=x0, maxiter=1, batched=True)
    assert popn.compute_log_p(x_fit) > lp0


def test_lock_step_hmc_targets_the_same_posterior_as_the_per_neuron_chains():
    """BatchedHmcGlmUpdate('bias') (all neurons share each engine call) and HmcBiasUpdate (the reference's
    neuron-by-neuron schedule, gibbs.py:164-321) are chains on the same conditional posterior of each bias:
    their sample means agree within Monte-Carlo error, and both sit at the optimum of the log posterior."""
    from theano_pyglm_b200.inference.gibbs import BatchedHmcGlmUpdate, HmcBiasUpdate
    model, popn, data, x_true = synth_standard_glm(N=3, nT=8000, seed=6)
    popn.add_data(data)
    N = 3

    def run(update_all, x, n_iter=300, burn=100):
        out = []
        for it in range(n_iter):
            update_all(x)
            if it >= burn:
                out.append([x['glms'][n]['bias']['bias'][0] for n in range(N)])
        return np.array(out)

    np.random.seed(0)
    xb = copy.deepcopy(x_true)
    ub = BatchedHmcGlmUpdate('bias', 10)
    ub.preprocess(popn)
    Sb = run(lambda x: ub.update(x), xb)
    np.random.seed(1)
    xs = copy.deepcopy(x_true)
    us = HmcBiasUpdate()
    us.preprocess(popn)
    Ss = run(lambda x: [us.update(x, n) for n in range(N)], xs, n_iter=160, burn=60)
    sd = np.maximum(Sb.std(axis=0), 1e-3)
    assert np.all(sd < 2.0)                                             # the data pin the bias down
    assert np.all(np.abs(Sb.mean(axis=0) - Ss.mean(axis=0)) < 1.0 * sd + 0.05)
    # the posterior mode of each bias (1-D Newton on the engine's gradient) lies inside the sampled cloud
    xm = copy.deepcopy(x_true)
    for n in range(N):
        for _ in range(30):
            lp, g = popn.glm_log_p_grad(xm, n)
            b0 = xm['glms'][n]['bias']['bias'][0]
            xm['glms'][n]['bias']['bias'] = np.array([b0 + 1e-3])
            g2 = (popn.glm_log_p_grad(xm, n)[1][0] - g[0]) / 1e-3
            xm['glms'][n]['bias']['bias'] = np.array([b0 - g[0] / g2])
        assert abs(xm['glms'][n]['bias']['bias'][0] - Sb.mean(axis=0)[n]) < 3.0 * sd[n] + 0.05


@pytest.mark.parametrize("kind", ["standard", "network", "stimulus"])
def test_dense_log_posterior_matches_the_per_neuron_dict_path(kind):
    """glms_log_p_grad_dense (priors, Dirichlet normalisation and chain rule vectorised over neurons) against
    glms_log_p_grad (component by component through the state dict) and glm_log_p_grad (one neuron)."""
    if kind == "standard":
        model, popn, data, x = synth_standard_glm(N=4, nT=5000, seed=3)
        popn.add_data(data)
    elif kind == "network":
        model, popn, data, x = synth_network_glm(N=12, nT=4000, seed=4)      # 12 > 10: 'g_10' sorts before 'g_2'
    else:
        N = 3
        model = make_model('standard_glm', N=N, dt=0.001)
        model['bkgd'] = {'type': 'basis', 'D_stim': 1, 'dt_max': 0.3, 'dt_stim': 0.1,
                         'basis': dict(type='cosine', n_eye=0, n_cos=3, a=1 / 120., b=0.5, orth=False, norm=True)}
        popn = Population(model)
        np.random.seed(2)
        x = popn.sample()
        for n in range(N):
            x['glms'][n]['imp']['w_ir'] *= 0.02
        S = (np.random.rand(3000, N) < 0.03).astype(float)
        popn.add_data({'S': S, 'N': N, 'dt': 0.001, 'T': 3.0, 'stim': np.random.randn(30, 1), 'dt_stim': 0.1})
    P = popn.dense_glm_params(x)
    lp_d, g_d = popn.glms_log_p_grad_dense(P, x)
    lp_s, g_s = popn.glms_log_p_grad(x)
    assert np.allclose(lp_d, lp_s, rtol=1e-12, atol=1e-9)
    assert np.allclose(g_d, g_s, rtol=1e-10, atol=1e-9)
    lp1, g1 = popn.glm_log_p_grad(x, 1)
    assert abs(lp1 - lp_d[1]) < 1e-9 * abs(lp1) and np.allclose(g1, g_d[1], rtol=1e-9, atol=1e-8)
    # round trip of the dense parameters
    x2 = copy.deepcopy(x)
    popn.set_dense_glm_params(x2, P * 1.5)
    assert np.allclose(popn.dense_glm_params(x2), P * 1.5)
