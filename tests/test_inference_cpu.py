"""CPU tests of the host-side samplers that replace the un-vendored `hips` (ARS, HMC) and of the
categorical draw rule."""
import numpy as np
import pytest

from oracle import pyglm_oracle as orc
from theano_pyglm_b200.inference.ars import adaptive_rejection_sample
from theano_pyglm_b200.inference.hmc import hmc
from theano_pyglm_b200.inference.log_sum_exp import log_sum_exp_sample


def test_log_sum_exp_sample_matches_oracle_rule():
    rng = np.random.default_rng(0)
    for _ in range(200):
        lnp = rng.standard_normal(int(rng.integers(2, 6))) * 5 + 700.0
        u = rng.random()
        assert log_sum_exp_sample(lnp, u) == orc.log_sum_exp_sample(lnp, u)
    np.random.seed(3)
    a = log_sum_exp_sample([0.0, 1.0])
    np.random.seed(3)
    assert a == orc.log_sum_exp_sample([0.0, 1.0], np.random.rand())      # consumes exactly one rand()
    with pytest.raises(Exception):
        log_sum_exp_sample([-np.inf, -np.inf], 0.5)


@pytest.mark.parametrize("mu,sig", [(0.0, 1.0), (-3.0, 0.2), (10.0, 5.0)])
def test_ars_samples_a_gaussian(mu, sig):
    np.random.seed(1)
    f = lambda w: -0.5 * ((w - mu) / sig) ** 2
    xs = np.sqrt(2) * sig * orc.GAUSS_HERMITE_ABSCISSAE + mu
    draws = np.array([adaptive_rejection_sample(f, xs, f(xs), (-np.inf, np.inf), stepsz=sig / 2) for _ in range(3000)])
    assert abs(draws.mean() - mu) < 5 * sig / np.sqrt(3000)
    assert abs(draws.std() - sig) < 0.06 * sig


def test_ars_skewed_log_concave_density_and_few_evaluations():
    """Poisson-like log posterior (the shape the W conditional has): mean by quadrature, and the number
    of func evaluations stays small (each one is a pass over the data on the GPU)."""
    np.random.seed(2)
    calls = [0]

    def f(w):
        calls[0] += 1
        return 3.0 * w - np.exp(w) - 0.5 * w ** 2
    xs = np.linspace(-2, 3, 10)
    n = 2000
    draws = np.array([adaptive_rejection_sample(f, xs, np.array([f(v) for v in xs]), (-np.inf, np.inf), stepsz=0.5)
                      for _ in range(n)])
    grid = np.linspace(-8, 6, 200001)
    dens = np.exp(3.0 * grid - np.exp(grid) - 0.5 * grid ** 2)
    mean = np.trapezoid(grid * dens, grid) / np.trapezoid(dens, grid)
    var = np.trapezoid((grid - mean) ** 2 * dens, grid) / np.trapezoid(dens, grid)
    assert abs(draws.mean() - mean) < 5 * np.sqrt(var / n)
    assert (calls[0] - 10 * n) / n < 3.0            # < 3 extra evaluations per draw beyond the given hull points


def test_ars_bounded_domain_and_sparse_start():
    np.random.seed(4)
    f = lambda w: -2.0 * w                           # exponential on [0, 5]
    draws = np.array([adaptive_rejection_sample(f, [0.5, 1.0, 2.0], [-1.0, -2.0, -4.0], (0.0, 5.0)) for _ in range(4000)])
    assert draws.min() >= 0.0 and draws.max() <= 5.0
    assert abs(draws.mean() - (0.5 - 5 * np.exp(-10) / (1 - np.exp(-10)))) < 0.03


def test_hmc_samples_a_gaussian_and_adapts():
    np.random.seed(5)
    mu, sig = np.array([1.0, -2.0]), np.array([0.5, 2.0])
    U = lambda q: 0.5 * np.sum(((q - mu) / sig) ** 2)
    gU = lambda q: (q - mu) / sig ** 2
    q, step, rate = np.zeros(2), 0.1, 0.9
    draws = []
    for _ in range(4000):
        q, step, rate = hmc(U, gU, step, 10, q, adaptive_step_sz=True, avg_accept_rate=rate)
        draws.append(q)
    d = np.array(draws[500:])
    assert np.all(np.abs(d.mean(axis=0) - mu) < 0.25 * sig)
    assert np.all(np.abs(d.std(axis=0) - sig) < 0.2 * sig)
    assert 1e-5 <= step <= 1.0 and 0.5 < rate <= 1.0


def test_hmc_batched_samples_independent_gaussians_and_respects_the_mask():
    """Lock-step HMC over M chains == M separate chains: right moments per chain, masked coordinates never move,
    chains without an active coordinate keep their step size and acceptance average."""
    from theano_pyglm_b200.inference.hmc import hmc_batched
    rng = np.random.RandomState(3)
    M, D = 5, 3
    mu = np.arange(M)[:, None] * np.ones((1, D))
    sig = np.array([0.5, 1.0, 2.0])[None, :]
    calls = [0]

    def U_and_grad(Q):
        calls[0] += 1
        z = (Q - mu) / sig
        return 0.5 * np.sum(z ** 2, axis=1), z / sig

    active = np.ones((M, D))
    active[1, 2] = 0.0                      # one frozen coordinate
    active[4, :] = 0.0                      # one frozen chain
    q = np.zeros((M, D))
    step, rate = np.full(M, 0.3), np.full(M, 0.9)
    samples = []
    for it in range(3000):
        q, step, rate = hmc_batched(U_and_grad, step, 10, q, active=active, avg_accept_rate=rate, rng=rng)
        if it >= 500:
            samples.append(q.copy())
    assert calls[0] == 3000 * 11            # one evaluation at the start point + one per leapfrog step, for ALL chains
    S = np.array(samples)
    free = active.astype(bool)
    assert np.max(np.abs(S.mean(axis=0) - mu)[free]) < 0.15
    assert np.max(np.abs(S.std(axis=0) / np.broadcast_to(sig, (M, D)) - 1.0)[free]) < 0.2      # 2500 correlated draws
    assert np.all(S[:, 1, 2] == 0.0) and np.all(S[:, 4, :] == 0.0)
    assert step[4] == 0.3 and rate[4] == 0.9 and np.all(step[:4] != 0.3)


def test_batched_edge_decisions_equal_the_per_edge_rule():
    """The vectorised decision of the lock-step sweep == _log_odds + log_sum_exp_sample edge by edge (gibbs.py:1002-1039,
    log_sum_exp.py:4-37), including NaN candidates, p_A = 0 / tiny / near 1 and uniforms on both sides of the threshold."""
    from theano_pyglm_b200.inference.gibbs import CollapsedGibbsNetworkColumnUpdate
    from theano_pyglm_b200.inference.log_sum_exp import log_sum_exp_sample
    upd = CollapsedGibbsNetworkColumnUpdate()
    rng = np.random.default_rng(5)
    M = 400
    ll = -1000.0 + 30.0 * rng.standard_normal((M, 11))
    ll[rng.random((M, 11)) < 0.02] = np.nan
    ll[:, 10] = np.where(np.isnan(ll[:, 10]), -1000.0, ll[:, 10])        # a NaN w=0 likelihood would make p(no edge) 0
    pA = rng.choice([0.003, 0.3, 0.5, 0.97], size=M)
    u = rng.random(M)
    a_vec = upd._decide_batch(ll, pA, u)

    for m in range(M):
        lp_no, lp_A = upd._log_odds(ll[m, :10], ll[m, 10], pA[m])
        lnp = np.array([lp_no, lp_A])
        p = np.exp(lnp - np.logaddexp(lp_no, lp_A))
        a_ref = 0 if u[m] <= p[0] else 1                                # first index with u <= cumsum (log_sum_exp.py:26-32)
        assert a_vec[m] == a_ref, (m, p, u[m])
    assert 0 < a_vec.sum() < M
