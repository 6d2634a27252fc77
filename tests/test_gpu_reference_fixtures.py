"""The CUDA engine, through the C ABI and the Population mirror, against outputs of THE REFERENCE ITSELF
(`tests/golden/ref_*.npz`, minted by `python -m oracle.ref_fixtures` from /root/reference)."""
import copy
import json
import os

import numpy as np
import pytest

import theano_pyglm_b200 as pg
from theano_pyglm_b200 import engine
from theano_pyglm_b200.inference.gibbs import CollapsedGibbsNetworkColumnUpdate
from theano_pyglm_b200.population import Population
from tests.test_oracle_vs_reference import POPS, _state_from_fixture, load, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(engine_lib):
    return engine_lib


def test_filter_kernels_reproduce_reference_convolve_with_basis(eng):
    """a1: K1 on the reference's own input/output pair of utils/basis.py:201-236 (integer counts up to 5, an arbitrary
    basis), and the dense FP64 filter on a real-valued stimulus."""
    f = load("ref_numpy_units.npz")
    ds = pg.Dataset(f["conv_S"], 0.001, f["conv_ibasis"], x_dtype="f64")
    assert rel(ds.fS(), f["conv_fS"]) < 1e-13
    ds.close()
    ds32 = pg.Dataset(f["conv_S"], 0.001, f["conv_ibasis"], x_dtype="f32")
    assert rel(ds32.fS(), f["conv_fS"]) < 1e-7          # FP32 storage of the FP64 sum
    ds32.close()
    got = engine.filter_dense(f["conv_stim"], f["conv_stim_ibasis"])
    assert rel(got, f["conv_fstim"]) < 1e-12


@pytest.mark.parametrize("path,tol_ll,tol_g", [("fp64", 1e-10, 1e-8), ("auto", 1e-6, 1e-5)])
@pytest.mark.parametrize("name", POPS)
def test_population_reproduces_reference_ll_prior_and_gradient(eng, name, path, tol_ll, tol_g):
    """a3-a11 through the reference-facing interface: Population(model).add_data(data); compute_ll / compute_log_prior /
    compute_log_p and the per-neuron (nlp, grad_nlp) of coord_descent.py:40-80 in the reference's packed order."""
    f = load(name)
    model = json.loads(str(f["model_json"]))
    N = model['N']
    popn = Population(model, path=path, x_dtype="f64" if path == "fp64" else "f32")      # "auto" + FP32 X = tcgen05 path
    x = _state_from_fixture(f, model)
    data = {'S': f["S"].astype(np.float64), 'N': N, 'dt': float(f["dt"]), 'T': float(f["T_sec"]),
            'stim': f["stim"] if "stim" in f else None, 'dt_stim': float(f["dt_stim"]) if "dt_stim" in f else None}
    popn.add_data(data)
    assert rel(popn.get_fS()[::50], f["fS_rows"]) < (1e-12 if popn.x_dtype == "f64" else 1e-7)
    if "fstim_rows" in f:
        assert rel(data['fstim'][::50], f["fstim_rows"]) < 1e-12
    ll = popn.ll_grad(x, grad=False)
    assert rel(ll, f["ll"]) < tol_ll
    assert abs(popn.compute_ll(x) - float(f["total_ll"])) < tol_ll * abs(float(f["total_ll"]))
    assert abs(popn.compute_log_prior(x) - float(f["total_log_prior"])) < 1e-10 * abs(float(f["total_log_prior"]))
    assert abs(popn.compute_log_p(x) - float(f["total_log_p"])) < tol_ll * abs(float(f["total_log_p"]))
    lam = popn.eval_state(x)['glms']
    assert rel(np.stack([lam[n]['lam'][::50] for n in range(N)], axis=1), f["lam_rows"]) < (1e-10 if popn.x_dtype == "f64" else 1e-7)
    lps, grads = popn.glms_log_p_grad(x)
    for n in range(N):
        assert np.array_equal(popn.glm_param_vector(x['glms'][n]), f["x_vec"][n])
        assert abs(-lps[n] - f["nlp"][n]) < tol_ll * abs(f["nlp"][n])
        assert rel(-grads[n], f["grad_nlp"][n]) < tol_g
        lp1, g1 = popn.glm_log_p_grad(x, n)
        assert abs(-lp1 - f["nlp"][n]) < tol_ll * abs(f["nlp"][n]) and rel(-g1, f["grad_nlp"][n]) < tol_g


def test_collapsed_gibbs_update_reproduces_reference_columns(eng, monkeypatch):
    """a12-a15: `CollapsedGibbsNetworkColumnUpdate.update(x, n)` of the mirror (K4 + host decision rule) replaying the
    random stream the reference consumed: same shuffled order, same 11 `_glm_ll` values per edge, identical A decisions
    and W values.  (The fixture generator replaced the un-vendored ARS by the posterior-mode grid point; so does this.)"""
    f = load("ref_sparse_weighted_n6.npz")
    model = json.loads(str(f["model_json"]))
    N = model['N']
    popn = Population(model)
    assert popn.x_dtype == "f64"
    x = _state_from_fixture(f, model)
    data = {'S': f["S"].astype(np.float64), 'N': N, 'dt': float(f["dt"]), 'T': float(f["T_sec"]), 'stim': None, 'dt_stim': None}
    popn.add_data(data)
    upd = CollapsedGibbsNetworkColumnUpdate()
    upd.preprocess(popn)

    def mode_point(ds, n_pre, n_post, mu_w, sigma_w, W_nns, log_L):
        lp = -0.5 / sigma_w ** 2 * (W_nns - mu_w) ** 2 + log_L
        ok = np.isfinite(lp) & (lp > -1e8)
        return W_nns[ok][np.argmax(lp[ok])]
    monkeypatch.setattr(upd, "_sample_w", mode_point)
    seen = []
    ds = popn._handle()
    real_delta = ds.gibbs_delta_ll

    def spy(cols, pres, w_cand):
        out = real_delta(cols, pres, w_cand)
        seen.append(out.copy())
        return out
    monkeypatch.setattr(ds, "gibbs_delta_ll", spy)
    for ci, n_post in enumerate(f["gibbs_cols"]):
        order = f["gibbs_order"][ci]
        us = iter([u for u in f["gibbs_uniforms"][ci] if not np.isnan(u)])
        zs = iter([z for z in f["gibbs_randn"][ci] if not np.isnan(z)])
        monkeypatch.setattr(np.random, "shuffle", lambda a: a.__setitem__(slice(None), order))
        monkeypatch.setattr(np.random, "rand", lambda *a: next(us))
        monkeypatch.setattr(np.random, "randn", lambda *a: next(zs))
        del seen[:]
        upd.begin(x)
        with np.errstate(divide='ignore'):
            upd.update(x, int(n_post))
        upd.end()
        ref_ll = f["gibbs_glm_ll"][ci]
        got = np.concatenate(seen, axis=0)
        assert got.shape == ref_ll.shape
        for i in range(N):
            fin = np.isfinite(ref_ll[i])
            near = fin & (ref_ll[i] > np.nanmax(ref_ll[i]) - 300.0)
            assert rel(got[i][near], ref_ll[i][near]) < 1e-10
            assert np.all(got[i][~fin] < np.nanmax(ref_ll[i]) - 100.0)      # NaN in the reference == "contributes nothing"
        assert np.array_equal(x['net']['graph']['A'], f["gibbs_A_after"][ci])
        assert rel(x['net']['weights']['W'].reshape(N, N), f["gibbs_W_after"][ci]) < 1e-12
