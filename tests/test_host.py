"""CPU tests of the host-side mirror of the reference interface (no GPU, no engine calls)."""
import numpy as np
import pytest

from oracle import pyglm_oracle as orc
from tests.helpers import make_ibasis
from theano_pyglm_b200.models.model_factory import check_stability, make_model, stabilize_sparsity
from theano_pyglm_b200.population import Population
from theano_pyglm_b200.utils import basis as pbasis
from theano_pyglm_b200.utils.packvec import get_vars, pack, packdict, set_vars, unpack, unpackdict
from theano_pyglm_b200.utils.parallel_util import neuron_shard, time_shard


def test_product_basis_matches_oracle_basis():
    for prms in (dict(type='cosine', n_eye=0, n_cos=5, a=1 / 120., b=0.5, orth=True, norm=False),
                 dict(type='cosine', n_eye=0, n_cos=7, a=1 / 120., b=0.5, orth=False, norm=True),
                 dict(type='cosine', n_eye=2, n_cos=3, a=1 / 120., b=0.5, orth=False, norm=False)):
        assert np.allclose(pbasis.create_basis(prms), orc.create_basis(prms), atol=1e-14)
    b = pbasis.create_basis(dict(type='cosine', n_eye=0, n_cos=5, a=1 / 120., b=0.5, orth=False, norm=True))
    assert np.allclose(pbasis.interpolate_basis(b, 0.001, 0.2, True, "linear"),
                       orc.interpolate_basis_linear(b, 0.001, 0.2, True))
    assert np.allclose(pbasis.interpolate_basis(b, 0.001, 0.2, True, "dirichlet"),
                       orc.interpolate_basis_dirichlet(b, 0.001, 0.2, True))
    assert np.array_equal(pbasis.make_standard_ibasis(5), make_ibasis(5))


def test_packvec_sorted_key_order_roundtrip():
    d = {'imp': {'w_ir': np.arange(6.0)}, 'bias': {'bias': np.array([7.0])}, 'bkgd': {}, 'nlin': {}}
    vec, shapes = packdict(d)
    assert np.array_equal(vec, np.concatenate([[7.0], np.arange(6.0)]))     # bias < imp (packvec.py:23)
    back = unpackdict(vec * 2, shapes)
    assert np.array_equal(back['imp']['w_ir'], 2 * np.arange(6.0)) and back['bias']['bias'][0] == 14.0
    v, shp = pack([np.ones((2, 3)), np.zeros(4)])
    parts = unpack(v, shp)
    assert parts[0].shape == (2, 3) and parts[1].shape == (4,)
    syms = {'bias': {'bias': None}}
    assert get_vars(syms, d) == {'bias': {'bias': d['bias']['bias']}}
    set_vars(syms, d, {'bias': {'bias': np.array([1.5])}})
    assert d['bias']['bias'][0] == 1.5


def test_population_state_layout_and_priors():
    model = make_model('sparse_weighted_model', N=5, dt=0.001)
    stabilize_sparsity(model)
    assert abs(model['network']['graph']['rho'] - min(1.0, (0.7 + 0.2) ** 2 / 5)) < 1e-12   # model_factory.py:90-102
    popn = Population(model)
    np.random.seed(0)
    x = popn.sample()
    assert set(x) == {'latent', 'net', 'glms'} and len(x['glms']) == 5
    assert x['net']['graph']['A'].dtype == np.int8 and x['net']['graph']['A'].shape == (5, 5)
    assert x['net']['weights']['W'].shape == (25,)
    assert sorted(x['glms'][2]['imp']) == ['g_%d' % i for i in range(5)] and x['glms'][2]['n'] == 2
    assert isinstance(check_stability(model, x, 5), bool)
    # priors against the oracle's restatement of the same formulas
    A, W = x['net']['graph']['A'], x['net']['weights']['W'].reshape(5, 5)
    rho = popn.network.graph.rho
    lp = orc.erdos_renyi_log_p(A, rho) + orc.gaussian_weight_log_p(W, 0.0, 1.0, -0.2, 0.5)
    for n in range(5):
        g = np.stack([x['glms'][n]['imp']['g_%d' % i] for i in range(5)])
        lp += orc.bias_log_prior(x['glms'][n]['bias']['bias'][0], 20.0, 0.25) + orc.dirichlet_impulse_log_p(g, 1)
    assert abs(popn.compute_log_prior(x) - lp) < 1e-9 * abs(lp)
    # engine parameter blocks: beta-normalised impulse weights, explicit A / W
    bias, w, A2, W2 = popn.glm.engine_params(x)
    assert w.shape == (5, 25) and np.allclose(w.reshape(5, 5, 5).sum(axis=2), 1.0)
    assert A2 is A and np.array_equal(W2, W)
    assert popn.x_dtype == "f64"                       # MCMC model -> FP64 filtered spike train


def test_standard_glm_population_and_group_lasso():
    model = make_model('standard_glm', N=4, dt=0.001)
    popn = Population(model)
    np.random.seed(1)
    x = popn.sample()
    assert x['net'] == {'graph': {}, 'weights': {}}
    assert x['glms'][0]['imp']['w_ir'].shape == (20,)
    lp = sum(orc.bias_log_prior(x['glms'][n]['bias']['bias'][0], 20, 0.1) +
             orc.group_lasso_log_p(x['glms'][n]['imp']['w_ir'].reshape(4, 5), 0.0, 10.0, 1.0) for n in range(4))
    assert abs(popn.compute_log_prior(x) - lp) < 1e-10 * abs(lp)
    gp = popn.glm.grad_log_prior(x['glms'][1])
    ref = orc.group_lasso_log_p_grad(x['glms'][1]['imp']['w_ir'].reshape(4, 5), 0.0, 10.0, 1.0)
    assert np.allclose(gp['imp']['w_ir'], ref.ravel())
    bias, w, A, W = popn.glm.engine_params(x)
    assert A is None and W is None and popn.x_dtype == "f32"
    vars_ = popn.get_variables()
    assert vars_['glm']['imp'] == {'w_ir': (20,)} and vars_['glm']['bias'] == {'bias': (1,)}
    assert popn.extract_vars(x, 2)['glm'] is x['glms'][2]
    with pytest.raises(Exception):
        make_model('no_such_model')


def test_engine_missing_library_fails_loudly(monkeypatch, tmp_path):
    from theano_pyglm_b200 import engine
    monkeypatch.setattr(engine, "_lib", None)
    monkeypatch.setattr(engine, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(engine.EngineError, match="no CPU fallback"):
        engine.load_library()


def test_spike_binning_is_integer_exact():
    from theano_pyglm_b200.engine import spikes_to_u8
    S = np.array([[0.0, 1.0], [3.0, 10.0]])
    out = spikes_to_u8(S)
    assert out.dtype == np.uint8 and np.array_equal(out, S.astype(np.uint8))
    for bad in (np.array([[0.5]]), np.array([[-1.0]]), np.array([[256.0]])):
        with pytest.raises(ValueError):
            spikes_to_u8(bad)


def test_shard_planners_cover_everything_once():
    for N, world in ((27, 8), (4, 8), (1024, 8), (5, 2)):
        spans = [neuron_shard(N, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    lo, hi, halo = time_shard(1000, 4, 0, 200)
    assert (lo, hi, halo) == (0, 250, 0)
    assert time_shard(1000, 4, 1, 200) == (250, 500, 200)
    assert time_shard(300, 4, 1, 200) == (75, 150, 75)      # halo clipped at the start of the recording


# ----------------------------------------------------------------------------------------------
# synthetic data, files, model conversion, stimulus component (host side, no GPU)
# ----------------------------------------------------------------------------------------------
class _GlobalStream:
    """Feeds oracle.simulate from numpy's global stream, the stream Population.simulate consumes."""

    @staticmethod
    def random(n):
        return np.random.rand(n)


def test_population_simulate_matches_oracle_and_reference_draw_order():
    N = 4
    model = make_model('standard_glm', N=N, dt=0.001)
    popn = Population(model)
    np.random.seed(11)
    x = popn.sample()
    for n in range(N):                                  # moderate coupling: a few hundred spikes in 3 s
        x['glms'][n]['imp']['w_ir'] *= 0.2
    np.random.seed(5)
    S, X = popn.simulate(x, (0, 3.0), 0.001, None, None)
    assert S.shape == (3000, N) and X.shape == (3000, N) and S.sum() > 50 and S.max() <= 10
    bias = np.array([x['glms'][n]['bias']['bias'][0] for n in range(N)])
    imps = np.stack([popn.glm.imp_model.impulse(x['glms'][n]['imp']) for n in range(N)], axis=1)
    np.random.seed(5)
    S_o, X_o = orc.simulate(bias, imps, np.ones((N, N), np.int8), np.ones((N, N)), 3000, 0.001, orc.NLIN_SOFTPLUS,
                            _GlobalStream)
    assert np.array_equal(S, S_o) and np.allclose(X, X_o, rtol=0, atol=1e-12)
    # the activation is the causal filter of the spikes: X = bias + fS . w  (generate_synth_data.py:124-129)
    fS = orc.convolve_with_basis(S, popn.glm.imp_model.ibasis)
    for n in range(N):
        act = bias[n] + orc.impulse_current(fS, x['glms'][n]['imp']['w_ir'].reshape(N, -1)).sum(axis=1)
        assert np.allclose(act, X[:, n], atol=1e-8)


def test_data_files_roundtrip_and_segments(tmp_path):
    from theano_pyglm_b200.utils import io as pio
    rng = np.random.default_rng(0)
    data = {'S': (rng.random((2000, 3)) < 0.05).astype(float), 'N': 3, 'dt': 0.001, 'T': 2.0,
            'stim': rng.standard_normal((20, 2)), 'dt_stim': 0.1, '_b200': object(), 'preprocessed': True}
    model = make_model('sparse_weighted_model', N=3)
    pio.save_results(str(tmp_path), data=data, model=model, results=[{'ll': 1.0}])
    back = pio.load_data(str(tmp_path / 'data.pkl'))
    assert '_b200' not in back and 'preprocessed' not in back
    assert np.array_equal(back['S'], data['S']) and back['N'] == 3 and back['dt_stim'] == 0.1
    assert pio.load_pickle(str(tmp_path / 'model.pkl')) == model
    assert pio.load_pickle(str(tmp_path / 'results.pkl')) == [{'ll': 1.0}]
    seg = pio.segment_data(data, (0.5, 1.5))
    # io.py:141-142 floors T // dt in floating point (0.5 // 0.001 == 499.0): same indices here
    i0, i1 = int(0.5 // 0.001), int(1.5 // 0.001)
    assert seg['T'] == 1.0 and seg['S'].shape == (i1 - i0, 3) and seg['stim'].shape == (10, 2)
    assert np.array_equal(seg['S'], data['S'][i0:i1]) and '_b200' not in seg
    with pytest.raises(Exception):
        pio.load_data(str(tmp_path / 'data.txt'))


def test_convert_model_projects_basis_impulses_onto_dirichlet_model():
    from theano_pyglm_b200.models.model_factory import convert_model
    N = 3
    m_std = make_model('standard_glm', N=N, dt=0.001)
    m_w = make_model('sparse_weighted_model', N=N, dt=0.001)
    p_std, p_w = Population(m_std), Population(m_w)
    np.random.seed(3)
    x_std, x_w = p_std.sample(), p_w.sample()
    # impulses that lie in the cone of the target basis, with known areas and signs
    ib_w = p_w.glm.imp_model.ibasis
    beta = np.random.dirichlet(np.ones(ib_w.shape[1]), size=(N, N))
    W_true = np.random.randn(N, N) * 2.0
    from scipy.linalg import lstsq
    for n2 in range(N):
        target = (W_true[:, n2, None] * beta[:, n2, :]) @ ib_w.T                     # (N_pre, R)
        w_ir = lstsq(p_std.glm.imp_model.ibasis, target.T)[0].T                      # best fit in the standard basis
        x_std['glms'][n2]['imp']['w_ir'] = w_ir.ravel()
    conv = convert_model(p_std, m_std, x_std, p_w, m_w, x_w)
    Wc = conv['net']['weights']['W'].reshape(N, N)
    assert conv['net']['graph']['A'].dtype == np.int8 and conv['net']['graph']['A'].shape == (N, N)
    for n2 in range(N):
        assert conv['glms'][n2]['bias']['bias'] is x_std['glms'][n2]['bias']['bias']
        imp_from = p_std.glm.imp_model.impulse(x_std['glms'][n2]['imp'])
        imp_to = p_w.glm.imp_model.impulse(conv['glms'][n2]['imp']) * Wc[:, n2, None]
        # the projection reproduces the impulse responses up to the fit residual of the two bases
        assert np.linalg.norm(imp_to - imp_from) < 0.25 * np.linalg.norm(imp_from)
        assert np.all(np.sign(Wc[:, n2]) == np.sign(W_true[:, n2]))


def test_basis_stimulus_component_matches_oracle():
    from theano_pyglm_b200.components.bkgd import BasisStimulus
    model = make_model('standard_glm', N=2, dt=0.001)
    model['bkgd'] = {'type': 'basis', 'D_stim': 2, 'dt_max': 0.3, 'dt_stim': 0.01,
                     'basis': dict(type='cosine', n_eye=0, n_cos=3, a=1 / 120., b=0.5, orth=False, norm=True)}
    comp = BasisStimulus(model)
    ib = orc.interpolate_stim_basis(orc.create_basis(model['bkgd']['basis']), 0.001, 0.3, True)
    assert np.allclose(comp.ibasis, ib, atol=1e-15) and comp.n_vars == 6
    assert np.allclose(comp.ibasis.sum(axis=0), 1.0)                                # sum-normalised (bkgd.py:119)
    w = np.linspace(-0.02, 0.03, 6)
    assert np.isclose(comp.log_p({'w_stim': w}), orc.stim_log_prior(w))
    eps = 1e-7
    g = comp.grad_log_p({'w_stim': w})['w_stim']
    for i in range(6):
        d = np.zeros(6); d[i] = eps
        fd = (comp.log_p({'w_stim': w + d}) - comp.log_p({'w_stim': w - d})) / (2 * eps)
        assert np.isclose(g[i], fd, rtol=1e-5)
    popn = Population(model)
    x = popn.sample()
    vec = popn.glm_param_vector(x['glms'][0])
    assert vec.size == 1 + 6 + 2 * 5                                                 # bias < bkgd < imp
    assert np.array_equal(vec[1:7], x['glms'][0]['bkgd']['w_stim'])
    popn.set_glm_param_vector(x['glms'][0], vec * 2)
    assert np.array_equal(x['glms'][0]['bkgd']['w_stim'], vec[1:7] * 2)
    assert popn.glm.stim_weights(x).shape == (2, 6)


def test_make_synth_dataset_packages_the_reference_data_dict():
    """generate_synth_data.py:56-135 without the files: model, prior draw (re-drawn until stable), simulation and
    the data dict the rest of the reference consumes."""
    from theano_pyglm_b200.utils.synth import make_synth_dataset
    model, popn, x_true, data = make_synth_dataset('standard_glm', N=3, T_stop=2.0, seed=4)
    assert set(['S', 'X', 'N', 'dt', 'T', 'stim', 'dt_stim', 'vars']) <= set(data)
    assert data['S'].shape == (2000, 3) and data['X'].shape == (2000, 3) and data['N'] == 3 and data['T'] == 2.0
    assert data['vars'] is x_true and data['stim'].shape == (20, 1)
    assert np.all(data['S'] == np.floor(data['S'])) and data['S'].min() >= 0 and data['S'].max() <= 10
    assert check_stability(model, x_true, 3)
    # the recorded activation reproduces the firing rate from bias + filtered spikes (generate_synth_data.py:124-129)
    fS = orc.convolve_with_basis(data['S'], popn.glm.imp_model.ibasis)
    for n in range(3):
        act = x_true['glms'][n]['bias']['bias'][0] + orc.impulse_current(fS, x_true['glms'][n]['imp']['w_ir'].reshape(3, -1)).sum(axis=1)
        assert np.allclose(act, data['X'][:, n], atol=1e-8)


def test_sta_and_stimulus_initialisation():
    """utils/sta.py:6-84 and smart_init.py:30-99 (temporal branch): the spike-triggered average against its definition,
    and the projected initial w_stim recovering a planted stimulus filter."""
    from theano_pyglm_b200.inference.smart_init import initialize_with_data
    from theano_pyglm_b200.models.model_factory import make_model
    from theano_pyglm_b200.population import Population
    from theano_pyglm_b200.utils.sta import project_onto_basis, sta
    rng = np.random.default_rng(3)
    nt, N, D, L = 400, 2, 2, 7
    data = {'S': (rng.random((nt, N)) < 0.2).astype(np.float64), 'dt': 0.001, 'dt_stim': 0.002}
    stim = rng.standard_normal((nt // 2, D))
    A = sta(stim, data, L, Ns=[1, 0])
    t = 0.001 * np.arange(nt)
    ist = np.stack([np.interp(t, 0.002 * np.arange(nt // 2), stim[:, d]) for d in range(D)], axis=1) / 2.0
    for i, n in enumerate([1, 0]):
        for l in range(L):
            ref = sum(data['S'][tt, n] * ist[tt - l] for tt in range(l, nt)) / data['S'][:, n].sum()
            assert np.allclose(A[i, l], ref, rtol=1e-12, atol=1e-14)
    basis = rng.standard_normal((L, 3))
    coef = rng.standard_normal(3)
    assert np.allclose(project_onto_basis(basis @ coef, basis).ravel(), coef)
    # smart_init on a BasisStimulus model: the STA of white noise driving a Poisson neuron points along its filter
    model = make_model('standard_glm', N=2, dt=0.001)
    model['bkgd'].update(type='basis', D_stim=1, dt_stim=0.001)
    popn = Population(model)
    ib = popn.glm.bkgd_model.ibasis
    nt = 60000
    stim = rng.standard_normal((nt, 1))
    w_true = np.array([[3.0, -1.0, 0.5], [-2.0, 2.0, 0.0]])
    drive = np.stack([np.convolve(stim[:, 0], np.concatenate([[0.0], ib @ w_true[n]]))[:nt] for n in range(2)], axis=1)
    S = (rng.random((nt, 2)) < 0.02 * np.exp(np.clip(drive, -3, 3))).astype(np.float64)
    data = {'S': S, 'dt': 0.001, 'dt_stim': 0.001, 'stim': stim, 'T': nt * 0.001}
    np.random.seed(0)
    x0 = popn.sample()
    x0['net']['graph']['A'] = np.zeros((2, 2), dtype=np.int8)
    initialize_with_data(popn, data, x0)
    assert np.all(x0['net']['graph']['A'] == 1)
    for n in range(2):
        w0 = x0['glms'][n]['bkgd']['w_stim']
        c = np.dot(ib @ w0, ib @ w_true[n]) / np.linalg.norm(ib @ w0) / np.linalg.norm(ib @ w_true[n])
        assert w0.shape == (3,) and c > 0.8


def test_fit_network_sets_gaussian_weights_to_the_prior_mean():
    from theano_pyglm_b200.inference.coord_descent import fit_network
    from theano_pyglm_b200.models.model_factory import make_model
    from theano_pyglm_b200.population import Population
    model = make_model('sparse_weighted_model', N=4, dt=0.001)
    popn = Population(model)
    np.random.seed(1)
    x = popn.sample()
    lp0 = popn.network.log_p(x['net'])
    fit_network(popn, x)
    W = x['net']['weights']['W'].reshape(4, 4)
    assert np.all(np.diag(W) == -0.2) and np.all(W[~np.eye(4, dtype=bool)] == 0.0)
    assert popn.network.log_p(x['net']) >= lp0
    x2 = Population(make_model('standard_glm', N=3, dt=0.001)).sample()
    assert fit_network(Population(make_model('standard_glm', N=3, dt=0.001)), x2) is x2      # constant weights: untouched


def test_shard_data_by_time_without_a_process_group_is_the_whole_recording():
    from theano_pyglm_b200.utils.parallel_util import shard_data_by_time, time_shard
    S = np.arange(40).reshape(20, 2)
    d = shard_data_by_time({'S': S, 'preprocessed': True, '_b200': object()}, R=7)
    assert d['halo'] == 0 and np.array_equal(d['S'], S) and '_b200' not in d and d['preprocessed'] is False
    # the planner the sharded form uses: contiguous blocks, left context of at most R bins
    assert [time_shard(20, 3, r, 7) for r in range(3)] == [(0, 7, 0), (7, 14, 7), (14, 20, 7)]
