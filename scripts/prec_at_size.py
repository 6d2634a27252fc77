"""Measure the tensor-core path against the float64 CPU oracle DIRECTLY at the benchmark shapes (C2 full size; C3, C4 and
N=4096 population sizes on recordings the oracle finishes in seconds): ll relative error, gradient error in max-norm and
element-wise over |g| > 1e-3 max|g|.  Prints one JSON line per case.   python scripts/prec_at_size.py [case ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyglm_oracle as orc          # noqa: E402
from tests.helpers import grad_errors, make_problem      # noqa: E402
import theano_pyglm_b200 as pg                  # noqa: E402

CASES = {"c2": (1_000_000, 27, 5, False), "c3": (100_000, 256, 5, True), "c4": (50_000, 1024, 10, True),
         "n4096": (8192, 4096, 5, True)}


def run(name, nlin=orc.NLIN_SOFTPLUS, x_dtype="f32"):
    T, N, B, network = CASES[name]
    if network:       # the benchmark's C3-style inputs: Dirichlet impulses, ER graph at the stabilised sparsity, Gaussian W
        from bench import make_gibbs_inputs
        g = make_gibbs_inputs(dict(N=N, T=T, B=B), 99)
        p = dict(S=g["S"], ibasis=g["ibasis"], bias=g["bias"], w=g["w"].reshape(N, N, B), A=g["A"], W=g["W"], dt=g["dt"])
    else:
        p = make_problem(T, N, B, seed=99)
    t0 = time.time()
    fS = orc.convolve_with_basis(p['S'].astype(np.float64), p['ibasis'])
    ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], p['bias'], p['w'], p['A'], p['W'], nlin)
    x = orc.population_activation(fS, p['bias'], p['w'], p['A'], p['W'])
    lam, _d, loglam = orc.nlin_and_derivative(x, nlin)
    terms = np.sum(np.abs(p['dt'] * lam) + np.abs(loglam * p['S']), axis=0)       # size of each neuron's sum
    t_cpu = time.time() - t0
    del fS, x, lam, loglam
    ds = pg.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype=x_dtype)
    out = {"case": name, "T": T, "N": N, "B": B, "cpu_s": round(t_cpu, 1), "path": ds.path_info("auto")["name"]}
    for path in ("tc", "fp64") if x_dtype == "f32" else ("tc",):
        if path == "fp64" and N * B > 6000:
            continue
        l, b, g = ds.ll_grad(p['bias'], p['w'].reshape(N, -1), p['A'], p['W'], nlin=nlin, path=path)
        gn = np.concatenate([gb[:, None], gw.reshape(N, -1)], axis=1)                 # each neuron's gradient vector
        gg = np.concatenate([b[:, None], g], axis=1)
        out[path] = {"ll_rel": float(np.max(np.abs(l - ll) / np.abs(ll))), "ll_total": float(abs(l.sum() - ll.sum()) / abs(ll.sum())),
                     "ll_vs_terms": float(np.max(np.abs(l - ll) / terms)),
                     "g_bias": grad_errors(b, gb), "g_w": grad_errors(g, gw.reshape(N, -1)),
                     "g_per_neuron_maxnorm": float(np.max(np.max(np.abs(gg - gn), axis=1) / np.max(np.abs(gn), axis=1)))}
    ds.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run(c)
