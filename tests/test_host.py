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
