"""TEST INFRASTRUCTURE -- mint golden vectors by RUNNING THE REFERENCE ITSELF.

    python -m oracle.ref_fixtures            (needs /root/reference; writes tests/golden/ref_*.npz)

`oracle/ref_loader.py` imports the unmodified sources under /root/reference/pyglm (mechanical
Python-2 -> 3 transform in memory, Theano replaced by `oracle/theano_shim.py`: the reference's own
graph-building code, evaluated literally in torch float64).  Every array written here is an output of
the reference's functions -- `Population.sample / simulate / add_data / compute_ll / compute_log_prior`,
`convolve_with_basis`, `prep_first_order_glm_inference` (`nlp`, `grad_nlp`),
`CollapsedGibbsNetworkColumnUpdate.update`, `log_sum_exp_sample`, `packdict`, `make_model`,
`stabilize_sparsity` -- on seeded inputs.  The fixtures travel to the GPU box; the reference does not.

What stays unpinned (stated in DESIGN.md section 2): Theano's own graph optimiser / reduction order
(the shim evaluates the expressions as written), and `hips` (ARS / HMC draws: un-vendored; the ARS
call is replaced here by a deterministic stand-in -- the posterior-mode grid point -- so that the `A`
decisions, which are fully specified in-tree, can be replayed).
"""
from __future__ import annotations

import copy
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def _quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o


# ---------------------------------------------------------------------------------------------
def fixture_numpy_units():
    """The reference's NumPy-only functions: basis.py, nlin.py (f_nlin), log_sum_exp.py, packvec.py,
    model_factory.py, models/*.py."""
    from pyglm.utils import basis as rb
    from pyglm.utils import packvec as rp
    from pyglm.inference import log_sum_exp as rl
    from pyglm.components import nlin as rn
    from pyglm.models.model_factory import make_model, stabilize_sparsity

    out = {}
    # a2: create_basis (utils/basis.py:9-106)
    for tag, prms in (("std5", dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=True, norm=False)),
                      ("std10", dict(type='cosine', n_eye=0, n_cos=10, a=1.0 / 120, b=0.5, orth=True, norm=False)),
                      ("dir5", dict(type='cosine', n_eye=0, n_cos=5, a=1.0 / 120, b=0.5, orth=False, norm=True)),
                      ("stim3", dict(type='cosine', n_eye=0, n_cos=3, a=1.0 / 120, b=0.5, orth=False, norm=True))):
        out["basis_" + tag] = rb.create_basis(prms)
    # a1: convolve_with_basis (utils/basis.py:201-236) on integer counts and on a real-valued signal
    rng = np.random.default_rng(21)
    S = (rng.random((1500, 3)) < 0.03).astype(np.float64)
    S[rng.integers(0, 1500, 12), rng.integers(0, 3, 12)] = rng.integers(2, 6, 12)
    ib = rng.standard_normal((200, 5)) * np.exp(-np.arange(200) / 60.0)[:, None]
    out["conv_S"], out["conv_ibasis"] = S.astype(np.uint8), ib
    out["conv_fS"] = rb.convolve_with_basis(S, ib)
    stim = np.cumsum(rng.standard_normal((700, 2)), axis=0) * 0.1
    ib2 = rng.standard_normal((37, 3))
    out["conv_stim"], out["conv_stim_ibasis"] = stim, ib2
    out["conv_fstim"] = rb.convolve_with_basis(stim, ib2)
    # a7: the numpy nonlinearities (components/nlin.py:29,47)
    xg = np.concatenate([np.linspace(-40, 40, 161), [-700.0, -100.0, 100.0, 700.0]])
    out["nlin_x"] = xg
    out["nlin_exp"] = rn.ExpNonlinearity({}).f_nlin(xg)
    with np.errstate(over='ignore'):
        out["nlin_explinear"] = rn.ExpLinearNonlinearity({}).f_nlin(xg)
    # a14: log_sum_exp_sample (inference/log_sum_exp.py:4-37) with np.random.rand replaced by a recorded stream
    cases = [np.array([0.0, 0.0]), np.array([-1000.0, -1001.0]), np.array([-5.0, 3.0, 1.0]),
             np.array([-np.inf, 2.0]), np.array([700.0, 710.0, -30.0]), np.array([1e-3, -1e-3])]
    us = [0.0, 0.25, 0.5, 0.7310585786300049, 0.9999999, 1.0 - 2 ** -53]
    lnps, uvals, choices = [], [], []
    real_rand = np.random.rand
    try:
        for lnp in cases:
            for u in us:
                np.random.rand = lambda *a, _u=u: _u
                try:
                    choices.append(rl.log_sum_exp_sample(lnp))
                except Exception:                      # "Invalid choice in logSumExp!" (:34-35): rounding left sum(p) < u
                    choices.append(-1)
                lnps.append(np.pad(lnp, (0, 3 - lnp.size), constant_values=np.nan))
                uvals.append(u)
    finally:
        np.random.rand = real_rand
    out["lse_lnp"], out["lse_u"], out["lse_choice"] = np.array(lnps), np.array(uvals), np.array(choices)
    # a9: sorted-key packing (utils/packvec.py:17-44, :62-84)
    nested = {'n': 3, 'bias': {'bias': np.array([20.5])}, 'bkgd': {'w_stim': np.arange(6.0) * 0.1},
              'imp': {'w_ir': np.arange(10.0) - 4.0}, 'nlin': {}}
    dn = {k: v for k, v in nested.items() if k != 'n'}
    vec, shapes = rp.packdict(dn)
    out["pack_vec"] = vec
    back = rp.unpackdict(vec * 2.0, shapes)
    out["pack_back_w_ir"] = back['imp']['w_ir']
    imp12 = {'g_%d' % i: np.full(2, float(i)) for i in range(12)}          # g_10 sorts before g_2
    out["pack_vec_g12"] = rp.packdict(imp12)[0]
    meta = {"pack_shapes": _jsonable(shapes)}
    # models (models/standard_glm.py, sparse_weighted_model.py, model_factory.py:18-102)
    meta["standard_glm_n4"] = _jsonable(make_model('standard_glm', N=4, dt=0.001))
    sw = make_model('sparse_weighted_model', N=256, dt=0.001)
    meta["sparse_weighted_n256_before"] = _jsonable(copy.deepcopy(sw))
    rhos = {}
    for N in (2, 6, 9, 27, 256, 1024):
        mN = make_model('sparse_weighted_model', N=N, dt=0.001)
        _quiet(stabilize_sparsity, mN)
        rhos[str(N)] = float(mN['network']['graph']['rho'])
    meta["stabilize_sparsity_rho"] = rhos
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "ref_numpy_units.npz"), **out)


# ---------------------------------------------------------------------------------------------
def _population_fixture(model, seed, T_sec, nlin_tag, fname, stim=None, dt_stim=None, gibbs_cols=()):
    """Run the reference end to end on `model`: prior sample, simulate, preprocess, ll / prior / gradient."""
    from pyglm.population import Population
    from pyglm.utils.theano_func_wrapper import seval
    from pyglm.utils.packvec import packdict, get_vars
    import pyglm.inference.coord_descent as rcd

    N, dt = model['N'], model['dt']
    popn = Population(model)
    np.random.seed(seed)
    x = popn.sample()
    nT = int(round(T_sec / dt))
    if stim is None:
        stim, dt_stim = np.zeros((nT, 1)), dt
        S, Xsim = _quiet(popn.simulate, x, (0, T_sec), dt, stim, dt_stim)
    else:
        # Population.simulate cannot run with a BasisStimulus (its temporary data dict has no 'T', bkgd.py:126):
        # spikes are seeded Bernoulli counts here and the simulator cross-check below is skipped
        S, Xsim = (np.random.rand(nT, N) < 0.03).astype(np.float64), None
    assert S.max() < 11 and S.sum() > 20 * N, "pick another seed: degenerate spike train (%s)" % S.sum(0)
    data = dict(S=S, N=N, dt=dt, T=T_sec, stim=stim, dt_stim=dt_stim)
    popn.add_data(data)
    syms = popn.get_variables()
    st = _quiet(popn.eval_state, x)
    # the reference's only numeric assertion on this path (test/generate_synth_data.py:124-129)
    for n in range(N if Xsim is not None else 0):
        assert np.allclose(st['glms'][n]['lam'], popn.glm.nlin_model.f_nlin(Xsim[:, n]))
    glm_syms, nlp, grad_nlp = _quiet(rcd.prep_first_order_glm_inference, popn)
    ll = np.zeros(N); lprior = np.zeros(N); nlps = np.zeros(N)
    vecs, grads = [], []
    for n in range(N):
        nv = popn.extract_vars(x, n)
        ll[n] = seval(popn.glm.ll, syms, nv)
        lprior[n] = seval(popn.glm.log_prior, syms, nv)
        vec, shapes = packdict(get_vars(glm_syms, nv['glm']))
        vecs.append(vec)
        nlps[n] = nlp(vec, nv)
        grads.append(grad_nlp(vec, nv))
    out = dict(S=S.astype(np.uint8), dt=dt, T_sec=T_sec, seed=seed,
               ibasis=popn.glm.imp_model.ibasis.get_value(), fS_rows=data['fS'][::50],
               Xsim_rows=Xsim[::50] if Xsim is not None else np.zeros(0), lam_rows=np.stack([st['glms'][n]['lam'][::50] for n in range(N)], axis=1),
               bias=np.array([x['glms'][n]['bias']['bias'][0] for n in range(N)]),
               ll=ll, log_prior_glm=lprior, nlp=nlps, x_vec=np.stack(vecs), grad_nlp=np.stack(grads),
               total_ll=popn.compute_ll(x), total_log_prior=popn.compute_log_prior(x), total_log_p=popn.compute_log_p(x),
               impulse=np.stack([st['glms'][n]['imp']['impulse'] for n in range(N)]),
               model_json=np.array(json.dumps(_jsonable(model))), shapes_json=np.array(json.dumps(_jsonable(shapes))),
               nlin=np.array(nlin_tag))
    if 'w_ir' in x['glms'][0]['imp']:
        out['w_ir'] = np.stack([x['glms'][n]['imp']['w_ir'] for n in range(N)])
    else:
        out['g'] = np.stack([np.stack([x['glms'][n]['imp']['g_%d' % k] for k in range(N)]) for n in range(N)])
    if 'w_stim' in x['glms'][0]['bkgd']:
        out['w_stim'] = np.stack([x['glms'][n]['bkgd']['w_stim'] for n in range(N)])
        out['stim'], out['dt_stim'] = stim, dt_stim
        out['fstim_rows'] = data['fstim'][::50]
        out['stim_ibasis'] = popn.glm.bkgd_model.ibasis.get_value()
    if 'A' in x['net']['graph']:
        out['A'] = x['net']['graph']['A'].copy()
        out['W'] = x['net']['weights']['W'].reshape(N, N).copy()
        out['p_A'] = seval(popn.network.graph.pA, syms['net'], x['net'])
        out['log_p_net'] = seval(popn.network.log_p, syms['net'], x['net'])
    if gibbs_cols:
        out.update(_gibbs_records(popn, x, gibbs_cols))
    np.savez_compressed(os.path.join(OUT, fname), **out)
    return out


def _gibbs_records(popn, x, cols):
    """CollapsedGibbsNetworkColumnUpdate.update (inference/gibbs.py:775-1250) on a copy of the state, with the
    random stream recorded and the un-vendored ARS replaced by the posterior-mode grid point."""
    import pyglm.inference.gibbs as rg
    import pyglm.inference.log_sum_exp as rl

    upd = rg.CollapsedGibbsNetworkColumnUpdate()
    upd.preprocess(popn)
    rg.adaptive_rejection_sample = lambda f, ws, lps, dom, stepsz=None, debug=False: ws[np.argmax(lps)]
    xg = copy.deepcopy(x)
    N = popn.N
    rec = dict(order=[], uniforms=[], randn=[], glm_ll=[], A_after=[], W_after=[])
    real_shuffle, real_rand, real_randn = np.random.shuffle, np.random.rand, np.random.randn
    real_glm_ll = upd._glm_ll

    def shuffle(a):
        real_shuffle(a)
        rec['order'].append(a.copy())

    def rand(*a):
        u = real_rand(*a)
        rec['uniforms'][-1].append(float(u))
        return u

    def randn(*a):
        z = real_randn(*a)
        rec['randn'][-1].append(float(z))
        return z

    def glm_ll(*a, **k):
        v = real_glm_ll(*a, **k)
        rec['glm_ll'][-1].append(float(v))
        return v

    upd._glm_ll = glm_ll
    np.random.shuffle, np.random.rand, np.random.randn = shuffle, rand, randn
    try:
        np.random.seed(99)
        with np.errstate(divide='ignore'):
            for n in cols:
                for k in ('uniforms', 'randn', 'glm_ll'):
                    rec[k].append([])
                upd.update(xg, n)
                rec['A_after'].append(xg['net']['graph']['A'].copy())
                rec['W_after'].append(xg['net']['weights']['W'].reshape(N, N).copy())
    finally:
        np.random.shuffle, np.random.rand, np.random.randn = real_shuffle, real_rand, real_randn
    pad = lambda rows, w: np.array([r + [np.nan] * (w - len(r)) for r in rows])
    return dict(gibbs_cols=np.array(cols), gibbs_order=np.stack(rec['order']), gibbs_uniforms=pad(rec['uniforms'], N),
                gibbs_randn=pad(rec['randn'], N), gibbs_glm_ll=np.array(rec['glm_ll']).reshape(len(cols), N, 11),
                gibbs_A_after=np.stack(rec['A_after']), gibbs_W_after=np.stack(rec['W_after']),
                gibbs_mu_w=np.array([upd.mu_w, upd.sigma_w, upd.mu_w_ref, upd.sigma_w_ref]))


def fixture_standard_glm():
    """C1's model (`generate_synth_data -m standard_glm -N 4`), 3 s of simulated spikes; explinear and exp."""
    from pyglm.models.model_factory import make_model
    m = make_model('standard_glm', N=4, dt=0.001)
    _population_fixture(m, seed=0, T_sec=3.0, nlin_tag="explinear", fname="ref_standard_glm_n4.npz")
    m2 = make_model('standard_glm', N=3, dt=0.001)
    m2['nonlinearity']['type'] = 'exp'
    m2['bias']['mu'] = 3.0
    m2['impulse']['prior']['sigma'] = 1.0
    _population_fixture(m2, seed=3, T_sec=3.0, nlin_tag="exp", fname="ref_standard_glm_n3_exp.npz")


def fixture_sparse_weighted():
    """C3's model (sparse_weighted_model: Dirichlet impulses, Erdos-Renyi graph, Gaussian weights), N=6, plus
    two collapsed-Gibbs column updates."""
    from pyglm.models.model_factory import make_model, stabilize_sparsity
    m = make_model('sparse_weighted_model', N=6, dt=0.001)
    _quiet(stabilize_sparsity, m)
    _population_fixture(m, seed=2, T_sec=2.0, nlin_tag="explinear", fname="ref_sparse_weighted_n6.npz", gibbs_cols=(2, 0))


def fixture_basis_stimulus():
    """standard_glm with a BasisStimulus background (components/bkgd.py:45-172)."""
    from pyglm.models.model_factory import make_model
    m = make_model('standard_glm', N=3, dt=0.001)
    m['bkgd']['type'] = 'basis'
    m['bkgd']['D_stim'] = 1          # the reference's own stim_response (bkgd.py:87) only evaluates for D_stim = 1
    m['bkgd']['dt_stim'] = 0.01
    rng = np.random.default_rng(8)
    T_sec = 2.0
    stim = np.cumsum(rng.standard_normal((int(T_sec / 0.01), 1)), axis=0) * 0.5
    _population_fixture(m, seed=4, T_sec=T_sec, nlin_tag="explinear", fname="ref_stimulus_glm_n3.npz", stim=stim, dt_stim=0.01)


def main():
    from oracle import ref_loader
    ref_loader.install()
    import scipy.integrate
    if not hasattr(scipy.integrate, 'cumtrapz'):           # renamed in SciPy since the reference was written
        scipy.integrate.cumtrapz = scipy.integrate.cumulative_trapezoid
    os.makedirs(OUT, exist_ok=True)
    fixture_numpy_units()
    fixture_standard_glm()
    fixture_sparse_weighted()
    fixture_basis_stimulus()
    for f in sorted(os.listdir(OUT)):
        if f.startswith("ref_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
