"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): 1e-6 relative on log-likelihood, 1e-5 on gradients;
integer work bit-exact.  The FP64 path is held to much tighter bounds.
"""
import numpy as np
import pytest

from oracle import pyglm_oracle as orc
from tests.helpers import make_problem, rel_err

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-6
GRAD_RTOL = 1e-5


@pytest.fixture(scope="module")
def eng(engine_lib):
    import theano_pyglm_b200 as pg
    return pg


def oracle_all(p, nlin):
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin)
    return fS, ll, gb, gw.reshape(p['N'], -1)


@pytest.mark.parametrize("T,N,B,x_dtype", [(3000, 4, 5, "f32"), (3000, 4, 5, "f64"), (5000, 27, 5, "f32"),
                                           (1111, 37, 10, "f32"), (700, 3, 1, "f64"), (2048, 70, 7, "f32")])
def test_filter_matches_oracle(eng, T, N, B, x_dtype):
    p = make_problem(T, N, B)
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype)
    fS = ds.fS()
    ref_fft = orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis'])      # as the reference calls it
    ref_dir = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    if x_dtype == "f64":
        assert np.array_equal(fS, ref_dir)            # same summation order -> bit-exact
    else:
        assert np.array_equal(fS, ref_dir.astype(np.float32).astype(np.float64))   # correctly rounded FP32
    assert np.max(np.abs(fS - ref_fft)) <= 1e-6 * max(1.0, np.max(np.abs(ref_fft)))
    ds.close()


def test_filter_long_window_and_heavy_counts(eng):
    """R >= 256 takes the runtime-pitch variant of K1; counts up to 255 in one bin; N a multiple of four above one
    column block (word-aligned row segments) and not (byte loads)."""
    from tests.helpers import make_ibasis
    for N, B in ((36, 5), (35, 3)):
        p = make_problem(1500, N, B)
        ib = make_ibasis(B, dt_max=0.3)
        assert ib.shape[0] >= 256
        S = p['S'].copy()
        S[[5, 700, 701, 1499], [0, N - 1, N - 1, 3]] = [255, 17, 2, 200]
        for x_dtype in ("f64", "f32"):
            ds = eng.Dataset(S, p['dt'], ib, x_dtype=x_dtype)
            ref = orc.convolve_with_basis_direct(S.astype(np.float64), ib)
            assert np.array_equal(ds.fS(), ref if x_dtype == "f64" else ref.astype(np.float32).astype(np.float64))
            ds.close()


def test_planes_written_by_the_filter_equal_planes_split_from_X(eng, monkeypatch):
    """Single-pass ingest: K1 writes the FP16 split planes itself (analytic per-feature scales).  The same planes built
    afterwards from the resident FP32 X (the path datasets with stimulus features take) give bit-identical results."""
    for T, N, B in ((3000, 27, 5), (2000, 70, 5), (1500, 9, 10)):
        p = make_problem(T, N, B, network=True)
        eager = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
        a = eager.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="tc")
        eager.refilter()                                   # a second pass rewrites X and planes in place
        a2 = eager.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="tc")
        eager.close()
        monkeypatch.setenv("PYGLM_LAZY_PLANES", "1")
        lazy = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
        monkeypatch.delenv("PYGLM_LAZY_PLANES")
        b = lazy.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="tc")
        lazy.close()
        for x, y, z in zip(a, b, a2):
            assert np.array_equal(x, y) and np.array_equal(x, z)


def test_filter_halo_matches_unsharded(eng):
    """Time-sharded ingest: shard k with an R-bin left halo reproduces rows of the unsharded X."""
    p = make_problem(4000, 6, 5)
    R = p['ibasis'].shape[0]
    full = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f64").fS()
    for lo, hi in [(0, 1000), (1000, 2500), (2500, 4000)]:
        halo = min(R, lo)
        ds = eng.Dataset(p['S'][lo - halo:hi], p['dt'], p['ibasis'], halo=halo, x_dtype="f64")
        assert np.array_equal(ds.fS(), full[lo:hi])
        ds.close()


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
@pytest.mark.parametrize("T,N,B,network", [(3000, 4, 5, False), (6000, 27, 5, False), (2500, 40, 10, True),
                                           (999, 5, 3, True)])
def test_ll_grad_fp64_path(eng, T, N, B, network, nlin):
    p = make_problem(T, N, B, network=network)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0          # exp(3) ~ 20 Hz
    _, ll, gb, gw = oracle_all(p, nlin)
    for x_dtype, tol in (("f64", 1e-11), ("f32", 2e-7)):
        ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype)
        ll_g, gb_g, gw_g = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="fp64")
        assert rel_err(ll_g, ll) < tol
        assert rel_err(gb_g, gb) < max(tol, 1e-9) * 10
        assert rel_err(gw_g, gw) < max(tol, 1e-9) * 10
        # ll-only call and a sub-range of neurons
        lo, hi = N // 3, N - 1
        ll_sub = ds.ll(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, n_lo=lo, n_hi=hi, path="fp64")
        assert rel_err(ll_sub, ll[lo:hi]) < tol
        ds.close()


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
@pytest.mark.parametrize("T,N,B,network", [(3000, 4, 5, False), (6000, 27, 5, False), (1111, 10, 10, True),
                                           (2500, 16, 10, True), (130, 5, 3, True), (40000, 27, 5, True),
                                           # N*B > 160: the two-kernel GEMM path (column blocks of 128, split-K over time)
                                           (3000, 40, 5, True), (5000, 64, 10, False), (2100, 130, 5, True),
                                           # C4's population size: 10240 features, 8 column blocks
                                           (1536, 1024, 10, True)])
def test_ll_grad_tensor_core_path(eng, T, N, B, network, nlin):
    """tcgen05 path (FP16 split planes, FP32 epilogue, FP64 sums) at the north-star tolerances:
    1e-6 relative on ll, 1e-5 on gradients, against the float64 oracle."""
    p = make_problem(T, N, B, network=network)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
    fS_o, ll, gb, gw = oracle_all(p, nlin)
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
    assert ds.path_info("auto")["name"].startswith("tcgen05")
    ll_g, gb_g, gw_g = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="tc")
    ll_den = np.abs(ll)
    if N * B > 2048:
        # 10^4-feature dot products accumulated in FP32: the bound is relative to the size of the sum's terms
        # (for the exp model at a 20 Hz rate, -dt*lam and S*log(lam) cancel to a small ll in some neurons)
        x = orc.population_activation(fS_o, p['bias'], p['w'], p['A'], p['W'])
        lam, _d, loglam = orc.nlin_and_derivative(x, nlin)
        ll_den = np.sum(np.abs(p['dt'] * lam) + np.abs(loglam * p['S']), axis=0)
    assert np.max(np.abs(ll_g - ll) / ll_den) < LL_RTOL, (ll_g, ll)
    assert rel_err(gb_g, gb) < GRAD_RTOL
    assert rel_err(gw_g, gw) < GRAD_RTOL
    # second call reuses the planes; sub-range of neurons; ll-only
    lo, hi = 1, N - 1
    ll_s, gb_s, gw_s = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, n_lo=lo, n_hi=hi, path="tc")
    assert np.max(np.abs(ll_s - ll[lo:hi]) / ll_den[lo:hi]) < LL_RTOL
    assert rel_err(gw_s, gw[lo:hi]) < GRAD_RTOL
    ll_only = ds.ll(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="tc")
    assert np.max(np.abs(ll_only - ll) / ll_den) < LL_RTOL
    # and it agrees with the exact FP64 path on the device
    ll_x, gb_x, gw_x = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="fp64")
    assert rel_err(gw_g, gw_x) < GRAD_RTOL and rel_err(gb_g, gb_x) < GRAD_RTOL
    ds.close()


@pytest.mark.skipif(__import__("os").environ.get("PYGLM_GEMM_EXPERIMENTS_BUILD") != "1",
                    reason="quarantined experiments: build with PYGLM_NVCC_EXTRA=-DPYGLM_GEMM_EXPERIMENTS=1 and set "
                           "PYGLM_GEMM_EXPERIMENTS_BUILD=1 to run them")
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_forward_cluster_variants(eng, mode, monkeypatch):
    """The forward GEMM kernel's other variants (32-feature chunks in 64-byte rows; X multicast over a pair of column
    blocks; cta_group::2 MMAs over 256 bins) stay off by default, but they must stay correct: same tolerances,
    shapes with even / odd numbers of time tiles and column blocks, several K segments."""
    monkeypatch.setenv("PYGLM_GEMM_MODE", str(mode))
    for (T, N, B) in ((2100, 130, 5), (900, 300, 8)):
        p = make_problem(T, N, B, network=True, seed=4)
        _, ll, gb, gw = oracle_all(p, orc.NLIN_SOFTPLUS)
        ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
        ll_g, gb_g, gw_g = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=orc.NLIN_SOFTPLUS, path="tc")
        assert np.max(np.abs(ll_g - ll) / np.abs(ll)) < LL_RTOL, (mode, T, N, B)
        assert rel_err(gb_g, gb) < GRAD_RTOL and rel_err(gw_g, gw) < GRAD_RTOL
        monkeypatch.delenv("PYGLM_GEMM_MODE")                          # the default variant
        ll_0 = ds.ll(p['bias'], p['w'], p['A'], p['W'], nlin=orc.NLIN_SOFTPLUS, path="tc")
        monkeypatch.setenv("PYGLM_GEMM_MODE", str(mode))
        assert np.max(np.abs(ll_0 - ll_g) / np.abs(ll)) < 2e-7          # same arithmetic, different tiling of the work
        ds.close()


def test_tensor_core_path_at_benchmark_size(eng):
    """BASELINE.json config C2 (N=27, T=1e6, B=5) is too big for the CPU oracle in a test, so the
    tensor-core path is held to the north-star tolerances against the device FP64 path (itself
    oracle-checked above), plus two size-independent properties: additivity of ll / gradients over
    time shards (glm.py:52 is a sum over t) and the exact ll of an all-zero parameter vector."""
    from bench import WORKLOADS, make_inputs
    wl = WORKLOADS["c2"]
    inp = make_inputs(wl, 1234)
    N, T = wl["N"], wl["T"]
    ds = eng.Dataset(inp["S"], inp["dt"], inp["ibasis"])
    ll_x, gb_x, gw_x = ds.ll_grad(inp["bias"], inp["w"], path="fp64")
    ll_t, gb_t, gw_t = ds.ll_grad(inp["bias"], inp["w"], path="tc")
    assert np.max(np.abs(ll_t - ll_x) / np.abs(ll_x)) < LL_RTOL
    assert rel_err(gb_t, gb_x) < GRAD_RTOL
    assert rel_err(gw_t, gw_x) < GRAD_RTOL
    # w = 0: activation is the bias, so ll has a closed form
    ll_0 = ds.ll(inp["bias"], np.zeros_like(inp["w"]), path="tc")
    lam = orc.nlin(inp["bias"], orc.NLIN_SOFTPLUS)
    nspk = inp["S"].sum(axis=0, dtype=np.float64)
    assert np.max(np.abs(ll_0 - (-inp["dt"] * lam * T + np.log(lam) * nspk)) / np.abs(ll_0)) < LL_RTOL
    ds.close()
    # time shards with an R-bin halo add up to the whole
    R = inp["ibasis"].shape[0]
    acc = [0.0, 0.0, 0.0]
    for lo, hi in [(0, 300_000), (300_000, 650_001), (650_001, T)]:
        halo = min(R, lo)
        sh = eng.Dataset(inp["S"][lo - halo:hi], inp["dt"], inp["ibasis"], halo=halo)
        out = sh.ll_grad(inp["bias"], inp["w"], path="tc")
        acc = [a + o for a, o in zip(acc, out)]
        sh.close()
    assert np.max(np.abs(acc[0] - ll_x) / np.abs(ll_x)) < LL_RTOL
    assert rel_err(acc[1], gb_x) < GRAD_RTOL and rel_err(acc[2], gw_x) < GRAD_RTOL


@pytest.mark.parametrize("T,N,B", [(3000, 27, 5), (2500, 70, 5)])
def test_planes_only_dataset(eng, T, N, B, monkeypatch):
    """x_dtype="planes": only the FP16 split planes are resident (K1 writes them directly, no FP32 X anywhere);
    results equal the regular dataset's bit for bit, and the entry points that need X itself refuse."""
    p = make_problem(T, N, B, network=True)
    _, ll, gb, gw = oracle_all(p, orc.NLIN_SOFTPLUS)
    full = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
    ref = full.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="tc")
    full.close()
    for _ in range(1):
        ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="planes")
        out = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="auto")
        for o, r in zip(out, ref):
            assert np.array_equal(o, r)
        assert np.max(np.abs(out[0] - ll) / np.abs(ll)) < LL_RTOL and rel_err(out[2], gw) < GRAD_RTOL
        with pytest.raises(eng.EngineError):
            ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="fp64")
        with pytest.raises(eng.EngineError):
            ds.fS()
        # the Gibbs entry points do work on a planes-only dataset: they gather their currents from the spike trains
        ds.gibbs_begin(p['bias'], p['w'], p['A'], p['W'])
        cand = np.linspace(-1.0, 1.0, 11)[None, :]
        got = ds.gibbs_delta_ll([1], [2], cand)
        ds.gibbs_end()
        f64 = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f64")
        f64.gibbs_begin(p['bias'], p['w'], p['A'], p['W'])
        assert np.allclose(got, f64.gibbs_delta_ll([1], [2], cand), rtol=1e-10, atol=1e-9)
        f64.gibbs_end()
        f64.close()
        ds.close()


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
@pytest.mark.parametrize("T,N,B,chunk", [(3000, 27, 5, None), (3000, 27, 5, "1024"), (2500, 70, 5, "640"),
                                         (1500, 9, 10, "512"), (700, 40, 10, "256")])
def test_ll_grad_from_spikes_equals_resident_planes(eng, T, N, B, chunk, nlin, monkeypatch):
    """From-spikes K2 (x_dtype="none"): every evaluation expands the spikes into chunk-sized operand planes (K1 with the
    dataset's analytic scales) and runs the tensor-core kernels chunk by chunk.  One chunk: bit-identical to the resident
    planes.  Several chunks: only the order of the final FP64 additions differs (1e-12), and the oracle bounds hold.
    Covers the fused kernel (135 / 90 features) and the GEMM path (350 / 400 features), sub-ranges of neurons (the
    neuron-sharded use), and a time shard with a left halo."""
    p = make_problem(T, N, B, network=True, seed=77)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
    _, ll, gb, gw = oracle_all(p, nlin)
    res = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="planes")
    ref = res.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="tc")
    lo, hi = N // 3, N - 2
    ref_sub = res.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="tc", n_lo=lo, n_hi=hi)
    res.close()
    if chunk:
        monkeypatch.setenv("PYGLM_STREAM_CHUNK", chunk)
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="none")
    for rep in range(3):                                   # the third call replays the captured CUDA graph
        out = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)
        for o, r in zip(out, ref):
            if chunk is None:
                assert np.array_equal(o, r)
            else:
                assert np.max(np.abs(o - r)) <= 1e-12 * max(1.0, np.max(np.abs(r)))
    assert np.max(np.abs(out[0] - ll) / np.maximum(np.abs(ll), 1e-3)) < LL_RTOL * max(1.0, (2000.0 / T) ** 0.5)
    assert rel_err(out[1], gb) < GRAD_RTOL and rel_err(out[2], gw) < GRAD_RTOL
    sub = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, n_lo=lo, n_hi=hi)
    for o, r in zip(sub, ref_sub):
        assert np.max(np.abs(o - r)) <= 1e-12 * max(1.0, np.max(np.abs(r)))
    assert np.allclose(ds.ll(p['bias'], p['w'], p['A'], p['W'], nlin=nlin), out[0], rtol=1e-13)
    ds.close()
    # a time shard with its left context: rows [cut, T) of the recording
    cut, R = T // 2 + 3, p['ibasis'].shape[0]
    halo = min(R, cut)
    a = eng.Dataset(p['S'][cut - halo:], p['dt'], p['ibasis'], halo=halo, x_dtype="planes")
    b = eng.Dataset(p['S'][cut - halo:], p['dt'], p['ibasis'], halo=halo, x_dtype="none")
    ra = a.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)
    rb = b.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)
    for o, r in zip(rb, ra):
        assert np.max(np.abs(o - r)) <= 1e-12 * max(1.0, np.max(np.abs(r)))
    a.close(); b.close()


def test_ll_grad_null_network_is_complete_graph(eng):
    p = make_problem(2000, 6, 5)
    _, ll, gb, gw = oracle_all(p, orc.NLIN_SOFTPLUS)
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f64")
    ll_g, gb_g, gw_g = ds.ll_grad(p['bias'], p['w'], None, None, path="fp64")
    assert rel_err(ll_g, ll) < 1e-11 and rel_err(gw_g, gw) < 1e-9
    ds.close()


def test_firing_rate_matches_simulated_activation(eng):
    """The reference's own assertion (test/generate_synth_data.py:124-129):
    lam from the likelihood graph == f_nlin(X) accumulated by Population.simulate."""
    rng = np.random.default_rng(7)
    N, B, nT, dt = 4, 5, 4000, 0.001
    from tests.helpers import make_ibasis
    ib = make_ibasis(B)
    bias = 20.0 + 0.1 * rng.standard_normal(N)
    w = orc.sample_group_lasso(rng, N * N, B, 0.0, 10.0, 1.0).reshape(N, N, B) * 0.02
    A = np.ones((N, N), dtype=np.int8)
    W = np.ones((N, N))
    imps = np.einsum('npb,rb->pnr', w, ib)             # (pre, post, R): impulse.py:65, population.py:281
    S, Xsim = orc.simulate(bias, imps, A, W, nT, dt, orc.NLIN_SOFTPLUS, rng)
    ds = eng.Dataset(S, dt, ib, x_dtype="f64")
    lam = ds.firing_rate(bias, w.reshape(N, -1), A, W)
    assert np.allclose(lam, orc.nlin(Xsim, orc.NLIN_SOFTPLUS))          # the reference's np.allclose
    assert rel_err(lam, orc.nlin(Xsim, orc.NLIN_SOFTPLUS)) < 1e-10
    ds.close()


def test_empty_and_tiny_inputs(eng):
    p = make_problem(50, 3, 5)
    # T = 0: empty recording -> zero ll and gradient
    ds = eng.Dataset(p['S'][:0], p['dt'], p['ibasis'])
    ll, gb, gw = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="fp64")
    assert np.all(ll == 0) and np.all(gb == 0) and np.all(gw == 0)
    assert ds.fS().shape == (0, 3, 5)
    ds.close()
    # T shorter than the filter, and a silent population
    for S in (p['S'], np.zeros_like(p['S'])):
        ds = eng.Dataset(S, p['dt'], p['ibasis'], x_dtype="f64")
        q = dict(p, S=S)
        fS, ll, gb, gw = oracle_all(q, orc.NLIN_SOFTPLUS)
        assert np.array_equal(ds.fS(), fS)
        ll_g, gb_g, gw_g = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], path="fp64")
        assert rel_err(ll_g, ll) < 1e-11
        assert np.allclose(gw_g, gw, rtol=1e-9, atol=1e-12)
        ds.close()
    # empty neuron range
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'])
    assert ds.ll(p['bias'], p['w'], n_lo=2, n_hi=2).shape == (0,)
    ds.close()


def test_bad_arguments_raise(eng):
    p = make_problem(100, 3, 5)
    with pytest.raises(ValueError):
        eng.Dataset(p['S'].astype(np.float64) + 0.5, p['dt'], p['ibasis'])
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'])
    with pytest.raises(eng.EngineError):
        ds.ll(p['bias'], p['w'], n_lo=2, n_hi=9)
    with pytest.raises(eng.EngineError):
        ds.gibbs_delta_ll([0], [1], np.zeros((1, 11)))       # before gibbs_begin
    ds.close()


@pytest.mark.parametrize("x_dtype", ["f64", "f32", "planes", "none"])
def test_gibbs_delta_ll_and_decisions(eng, x_dtype):
    """K4 vs gibbs.py:910-937 / :1002-1039 restated: candidate log-likelihoods, then identical
    A decisions for the same uniforms, through a whole shuffled column sweep with commits."""
    T, N, B = 4000, 6, 5
    p = make_problem(T, N, B, network=True, dirichlet=True, seed=99)
    rng = np.random.default_rng(5)
    mu_w, sig_w, mu_ref, sig_ref = 0.0, 1.0, -0.2, 0.5
    p_A = np.full((N, N), 0.5)
    np.fill_diagonal(p_A, 1.0 - 1e-3)
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype)
    A_ref, W_ref = p['A'].copy(), p['W'].copy()
    ds.gibbs_begin(p['bias'], p['w'], A_ref, W_ref, nlin="explinear")
    # FP32 storage of X perturbs u = X.w at ~1e-8 relative; candidate lls are differences of O(1e4)
    # sums, so the FP32-X bound is absolute.  Decision parity is asserted for both; the sampler
    # itself uses FP64 X (Population picks x_dtype="f64" for MCMC).
    # "planes" / "none" keep no X at all: K4 gathers its currents from the spike trains in FP64 (from-spikes mode)
    rtol, atol = (1e-7, 2e-5) if x_dtype == "f32" else (1e-10, 1e-9)
    for n_post in (1, 4):
        order = rng.permutation(N)
        unif = rng.random(N)
        wnew = rng.standard_normal(N)

        def w_draw(n_pre, a_new, mu, sig, W_nns, log_L, _w=wnew):
            return mu + sig * _w[n_pre]
        rec = orc.collapsed_column_sweep(fS, p['S'], p['dt'], n_post, p['bias'][n_post], p['w'][n_post], A_ref, W_ref,
                                         p_A, orc.NLIN_SOFTPLUS, mu_w, sig_w, mu_ref, sig_ref, order, unif, w_draw)
        for i, n_pre in enumerate(order):
            mu, sig = (mu_ref, sig_ref) if n_pre == n_post else (mu_w, sig_w)
            cand = np.concatenate([orc.gh_candidates(mu, sig), [0.0]])
            out = ds.gibbs_delta_ll([n_post], [n_pre], cand[None, :])[0]
            assert np.allclose(out[:10], rec[i]['log_L'], rtol=rtol, atol=atol)
            assert abs(out[10] - rec[i]['ll_noA']) <= rtol * abs(rec[i]['ll_noA']) + atol
            lp_noA, lp_A = orc.collapsed_edge_log_odds(out[:10], out[10], p_A[n_pre, n_post])
            a_new = orc.log_sum_exp_sample([lp_noA, lp_A], unif[i])
            assert a_new == rec[i]['A']                      # identical accept / reject
            ds.gibbs_commit([n_post], [n_pre], [a_new], [mu + sig * wnew[n_pre]])
    A_g, W_g = ds.gibbs_state()
    assert np.array_equal(A_g, A_ref)                            # integer state bit-exact
    assert np.array_equal(W_g, W_ref)
    # batched call over distinct columns == one-by-one calls
    cols = np.array([0, 2, 3, 5]); pres = np.array([3, 3, 0, 5])
    cand = rng.standard_normal((4, 11))
    batch = ds.gibbs_delta_ll(cols, pres, cand)
    for m in range(4):
        one = ds.gibbs_delta_ll(cols[m:m + 1], pres[m:m + 1], cand[m:m + 1])
        assert np.array_equal(batch[m], one[0])
    ds.gibbs_end()
    ds.close()


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
def test_gibbs_from_spikes_equals_the_filtered_spike_train_form(eng, nlin):
    """From-spikes mode of K4 (u[t] = sum_k (ibasis . w)[k] S[t-k], X never read) against the X-based mode on the same
    edges: several 8192-bin chunks (windows reach into the previous chunk), multi-spike bins, B = 10, a time shard with
    a left halo, batched edges, and commits (rank-1 updates of the resident I_net) in between."""
    T, N, B = 20011, 7, 10
    p = make_problem(T, N, B, network=True, dirichlet=True, seed=31, rate=0.03)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 18.0
        p['W'] = p['W'] * 0.05
    rng = np.random.default_rng(2)
    for lo, halo in ((0, 0), (5003, 200), (5003, 57)):
        S = p['S'][lo - halo:]
        ds_x = eng.Dataset(S, p['dt'], p['ibasis'], halo=halo, x_dtype="f64")
        ds_s = eng.Dataset(S, p['dt'], p['ibasis'], halo=halo, x_dtype="none")
        with pytest.raises(eng.EngineError):
            ds_s.ll(p['bias'], p['w'], p['A'], p['W'], path="fp64")    # spikes-only: no filtered spike train to read
        for ds in (ds_x, ds_s):
            ds.gibbs_begin(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)
        for rnd in range(3):
            cols = rng.permutation(N)[:5]
            pres = rng.integers(0, N, size=5)
            cand = rng.standard_normal((5, 11)) * (0.05 if nlin == orc.NLIN_EXP else 1.0)
            a, b = ds_x.gibbs_delta_ll(cols, pres, cand), ds_s.gibbs_delta_ll(cols, pres, cand)
            assert np.allclose(a, b, rtol=1e-10, atol=1e-9), (lo, halo, rnd, np.max(np.abs(a - b)))
            a_new = rng.integers(0, 2, size=5).astype(np.int8)
            w_new = rng.standard_normal(5) * (0.05 if nlin == orc.NLIN_EXP else 1.0)
            for ds in (ds_x, ds_s):
                ds.gibbs_commit(cols, pres, a_new, w_new)
        Ax, Wx = ds_x.gibbs_state()
        As, Ws = ds_s.gibbs_state()
        assert np.array_equal(Ax, As) and np.array_equal(Wx, Ws)
        for ds in (ds_x, ds_s):
            ds.gibbs_end()
            ds.close()


# ----------------------------------------------------------------------------------------------
# Stimulus features (BasisStimulus, bkgd.py:45-172)
# ----------------------------------------------------------------------------------------------
def make_stimulus(T, D, Bs, seed, dt=0.001):
    """A slowly varying D-dimensional stimulus sampled at 10 ms, filtered the way bkgd.py:122-154 does."""
    rng = np.random.default_rng(seed)
    dt_stim = 0.01
    stim = np.cumsum(rng.standard_normal((int(np.ceil(T * dt / dt_stim)) + 1, D)), axis=0) * 0.1
    prms = dict(type='cosine', n_eye=0, n_cos=Bs, a=1.0 / 120, b=0.5, orth=False, norm=True)
    ib = orc.interpolate_stim_basis(orc.create_basis(prms), dt, 0.3, True)
    istim, fstim = orc.filter_stimulus(stim, dt_stim, T, dt, ib)
    return istim, ib, fstim


def test_dense_filter_matches_oracle(eng):
    """engine.filter_dense vs convolve_with_basis (utils/basis.py:201-236) on a real-valued signal."""
    istim, ib, fstim = make_stimulus(5000, 2, 4, seed=3)
    out = eng.engine.filter_dense(istim, ib)
    assert out.shape == (5000, 2, 4)
    scale = np.max(np.abs(fstim))
    assert np.max(np.abs(out.reshape(5000, -1) - fstim)) < 1e-12 * scale
    # first bin has no past: exactly zero (zero row prepended, basis.py:220)
    assert np.all(out[0] == 0.0)


@pytest.mark.parametrize("nlin", [orc.NLIN_SOFTPLUS, orc.NLIN_EXP])
@pytest.mark.parametrize("T,N,B,D,Bs,network", [(3000, 6, 5, 2, 4, False),      # 38 features: fused kernel
                                                (4000, 27, 5, 1, 3, True),      # 138 features: fused kernel
                                                (2500, 40, 5, 2, 3, True)])     # 206 features: GEMM path
def test_ll_grad_with_stimulus(eng, T, N, B, D, Bs, network, nlin):
    """I_stim = fstim @ w_stim (bkgd.py:81) rides in the feature matrix: ll, d/d bias, d/d w_ir and
    d/d w_stim on the FP64 path and the tensor-core path against the oracle."""
    p = make_problem(T, N, B, network=network, seed=21)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
    _, _, fstim = make_stimulus(T, D, Bs, seed=8)
    F = D * Bs
    w_stim = 0.05 * np.random.default_rng(4).standard_normal((N, F))
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ll, gb, gw, gs = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin,
                                            fstim=fstim, w_stim=w_stim)
    ll0 = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin)[0]
    assert np.max(np.abs(ll - ll0) / np.abs(ll0)) > 1e-4        # the stimulus term matters in this problem
    gw = gw.reshape(N, -1)
    for x_dtype, path, lt, gt in (("f64", "fp64", 1e-11, 1e-9), ("f32", "fp64", LL_RTOL, GRAD_RTOL),
                                  ("f32", "tc", LL_RTOL, GRAD_RTOL)):
        ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype, fstim=fstim)
        assert ds.F == F
        ll_g, gb_g, gw_g, gs_g = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path=path, w_stim=w_stim)
        assert np.max(np.abs(ll_g - ll) / np.abs(ll)) < lt, (x_dtype, path)
        assert rel_err(gb_g, gb) < gt and rel_err(gw_g, gw) < gt and rel_err(gs_g, gs) < gt, (x_dtype, path)
        # neuron sub-range and ll-only
        ll_s, _, gw_s, gs_s = ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, n_lo=2, n_hi=N - 1, path=path,
                                         w_stim=w_stim)
        assert np.max(np.abs(ll_s - ll[2:N - 1]) / np.abs(ll[2:N - 1])) < lt
        assert rel_err(gs_s, gs[2:N - 1]) < gt and rel_err(gw_s, gw[2:N - 1]) < gt
        assert np.max(np.abs(ds.ll(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path=path, w_stim=w_stim) - ll)
                      / np.abs(ll)) < lt
        with pytest.raises(ValueError):
            ds.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)          # w_stim is required
        assert np.array_equal(ds.fS().reshape(T, -1), eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype).fS().reshape(T, -1))
        ds.close()


def test_gibbs_delta_ll_with_stimulus_and_extreme_activations(eng):
    """K4 with a stimulus current in the base activation (gibbs.py:914: I_bias + I_stim + I_net), and with
    weights large enough that candidates land in every branch of the tabulated softplus (|x| > 37 on both
    sides, spike bins at negative activation)."""
    T, N, B = 6000, 5, 5
    p = make_problem(T, N, B, network=True, dirichlet=True, seed=31)
    p['W'] = p['W'] * 0.3
    _, _, fstim = make_stimulus(T, 1, 3, seed=12)
    w_stim = 2.0 * np.random.default_rng(2).standard_normal((N, 3))
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f64", fstim=fstim)
    ds.gibbs_begin(p['bias'], p['w'], p['A'], p['W'], nlin="explinear", w_stim=w_stim)
    rng = np.random.default_rng(0)
    seen_lo = seen_hi = False
    for n_post, n_pre in ((0, 3), (2, 2), (4, 1)):
        I_imp = orc.impulse_current(fS, p['w'][n_post])
        Weff = orc.effective_weights(p['A'], p['W'], n_post).copy()
        Weff[n_pre] = 0.0
        I_other = I_imp @ Weff
        I_stim = fstim @ w_stim[n_post]
        cand = np.concatenate([orc.gh_candidates(0.0, 0.5), [0.0]])
        out = ds.gibbs_delta_ll([n_post], [n_pre], cand[None, :])[0]
        for q, wq in enumerate(cand):
            ref = orc.gibbs_glm_ll(p['bias'][n_post], I_stim, I_other, I_imp[:, n_pre], wq, p['S'][:, n_post], p['dt'],
                                   orc.NLIN_SOFTPLUS)
            assert abs(out[q] - ref) <= 1e-10 * abs(ref) + 1e-9, (n_post, n_pre, q, out[q], ref)
            x = p['bias'][n_post] + I_stim + I_other + wq * I_imp[:, n_pre]
            seen_lo |= bool(np.any(x < -37.0)); seen_hi |= bool(np.any(x > 37.0))
        ds.gibbs_commit([n_post], [n_pre], [1], [rng.standard_normal() * 0.5])
        p['A'][n_pre, n_post] = 1
        p['W'][n_pre, n_post] = ds.gibbs_state()[1][n_pre, n_post]
    assert seen_lo and seen_hi                                   # the test really visited both tails
    ds.gibbs_end()
    ds.close()


# ----------------------------------------------------------------------------------------------
# Randomised sweep over ragged shapes
# ----------------------------------------------------------------------------------------------
def _ragged_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for i in range(28):
        T = int(rng.choice([1, 2, 31, 127, 128, 129, 255, 257, 700, 1500]))
        N = int(rng.choice([1, 2, 3, 7, 27, 31, 32, 33, 45]))
        B = int(rng.choice([1, 2, 5, 8, 10, 16]))
        R = int(rng.choice([1, 3, 50, 200]))
        cases.append((i, T, N, B, R, bool(rng.integers(2)), int(rng.integers(2)), float(rng.choice([0.0, 0.02, 0.3]))))
    return cases


@pytest.mark.parametrize("case", _ragged_cases(), ids=lambda c: "c%d-T%d-N%d-B%d-R%d" % c[:5])
def test_ragged_shapes_all_paths(eng, case):
    """Shapes that do not fill a tile, a warp, a chunk or a basis window (T < R, N = 1, B = 16, silent and
    very active recordings): filter bit-exact in float64, ll / gradients on every path that applies."""
    i, T, N, B, R, network, nlin, rate = case
    p = make_problem(T, N, B, seed=500 + i, rate=rate, network=network, R=R)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
    S = p['S'].astype(np.float64)
    fS = orc.convolve_with_basis_direct(S, p['ibasis'])
    ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin)
    gw = gw.reshape(N, -1)
    ll_den = np.maximum(np.abs(ll), 1e-3)                         # a silent one-bin recording has ll ~ -dt*lam ~ 0.02
    d64 = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f64")
    assert np.array_equal(d64.fS(), fS)                            # same summation order as the oracle's causal sum
    l, b, w = d64.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path="fp64")
    assert np.max(np.abs(l - ll) / ll_den) < 1e-11 and rel_err(b, gb) < 1e-9
    assert np.max(np.abs(w - gw)) <= 1e-9 * max(np.max(np.abs(gw)), 1e-12)
    d64.close()
    d32 = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
    # The FP32 epilogue rounds every bin's term at ~2^-23 relative; over the long recordings of the configs these
    # roundings average out (1e-7 at T = 1e6), over a hundred bins they do not, so the bound widens as 1/sqrt(T).
    tc_tol = LL_RTOL * max(1.0, (2000.0 / T) ** 0.5)
    for path in ("fp64", "tc"):
        l, b, w = d32.ll_grad(p['bias'], p['w'], p['A'], p['W'], nlin=nlin, path=path)
        assert np.max(np.abs(l - ll) / ll_den) < (tc_tol if path == "tc" else LL_RTOL), (path, l, ll)
        assert np.max(np.abs(b - gb)) <= GRAD_RTOL * max(np.max(np.abs(gb)), 1e-6), path
        assert np.max(np.abs(w - gw)) <= GRAD_RTOL * max(np.max(np.abs(gw)), 1e-6), path
    d32.close()


def _gibbs_cases():
    rng = np.random.default_rng(77)
    out = []
    for i in range(14):
        out.append((i, int(rng.choice([1, 37, 255, 256, 257, 1000, 8192, 8193, 9000])), int(rng.choice([1, 2, 5, 9])),
                    int(rng.choice([1, 5, 10, 16])), int(rng.choice([1, 3, 4, 11, 16])), int(rng.integers(2)),
                    str(rng.choice(["f64", "f32"]))))
    return out


@pytest.mark.parametrize("case", _gibbs_cases(), ids=lambda c: "g%d-T%d-N%d-B%d-Q%d-nl%d-%s" % c)
def test_gibbs_delta_ll_ragged(eng, case):
    """K4 on shapes that do not fill a block (8192 bins), a 256-bin pass or a warp, for every candidate-count
    and basis-size template, both nonlinearities: each candidate ll against gibbs.py:910-937 restated."""
    i, T, N, B, Q, nlin, x_dtype = case
    p = make_problem(T, N, B, seed=900 + i, network=True, dirichlet=bool(i % 2), rate=0.05)
    if nlin == orc.NLIN_EXP:
        p['bias'] = p['bias'] - 17.0
        p['W'] = p['W'] * (0.02 if i % 2 else 1.0)                  # unit-area impulses under exp: keep rates finite
    rng = np.random.default_rng(i)
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype)
    ds.gibbs_begin(p['bias'], p['w'], p['A'], p['W'], nlin=nlin)
    n_post, n_pre = int(rng.integers(N)), int(rng.integers(N))
    cand = rng.standard_normal(Q) * (0.02 if (nlin == orc.NLIN_EXP and i % 2) else 1.0)
    out = ds.gibbs_delta_ll([n_post], [n_pre], cand[None, :])[0]
    I_imp = orc.impulse_current(fS, p['w'][n_post])
    Weff = orc.effective_weights(p['A'], p['W'], n_post).copy()
    Weff[n_pre] = 0.0
    I_other = I_imp @ Weff
    for q in range(Q):
        ref = orc.gibbs_glm_ll(p['bias'][n_post], 0.0, I_other, I_imp[:, n_pre], cand[q], p['S'][:, n_post], p['dt'], nlin)
        if not np.isfinite(ref):
            assert not np.isfinite(out[q]) or abs(out[q]) > 1e300
            continue
        scale = abs(ref) + 1e-3
        tol = (1e-10 if x_dtype == "f64" else 3e-6) * scale + (1e-9 if x_dtype == "f64" else 2e-5)
        assert abs(out[q] - ref) <= tol, (q, out[q], ref)
    ds.gibbs_end()
    ds.close()


def test_host_entry_graph_replay_sees_fresh_parameters(eng):
    """pyglm_b200_ll_grad replays a captured CUDA graph from its third call with the same signature on: every call
    must still read the caller's current parameter values, survive interleaved calls that move engine buffers,
    and match the oracle."""
    p = make_problem(4000, 27, 5, network=True, seed=8)
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    ds = eng.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
    rng = np.random.default_rng(0)
    for it in range(7):
        bias = p['bias'] + 0.05 * rng.standard_normal(27)
        w = p['w'] * (1.0 + 0.1 * it)
        W = p['W'] + 0.01 * it
        ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], bias, w, p['A'], W, orc.NLIN_SOFTPLUS)
        ll_g, gb_g, gw_g = ds.ll_grad(bias, w, p['A'], W, nlin="explinear", path="tc")
        assert np.max(np.abs(ll_g - ll) / np.abs(ll)) < LL_RTOL, it
        assert rel_err(gb_g, gb) < GRAD_RTOL and rel_err(gw_g, gw.reshape(27, -1)) < GRAD_RTOL, it
        if it == 3:                                 # other entry points grow workspaces between replays
            ds.firing_rate(bias, w, p['A'], W)
            ds.ll_grad(bias, w, p['A'], W, nlin="explinear", path="fp64", n_lo=3, n_hi=20)
        if it == 5:
            assert np.allclose(ds.ll(bias, w, p['A'], W, nlin="explinear", path="tc"), ll_g, rtol=1e-12)
    ds.close()
