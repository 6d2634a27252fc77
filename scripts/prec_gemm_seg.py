"""Precision of the GEMM-path forward accumulation vs the TMEM segment length (PYGLM_GEMM_SEG), N=1024 B=10."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theano_pyglm_b200 as pg
from oracle import pyglm_oracle as orc
from tests.helpers import make_problem, rel_err

for (T, N, B) in ((1536, 1024, 10), (3000, 256, 5)):
    p = make_problem(T, N, B, network=True)
    fS = orc.convolve_with_basis_direct(p['S'].astype(np.float64), p['ibasis'])
    for nlin, shift in ((orc.NLIN_SOFTPLUS, 0.0), (orc.NLIN_EXP, -17.0)):
        bias = p['bias'] + shift
        ll, gb, gw = orc.population_ll_grad(fS, p['S'], p['dt'], bias, p['w'], p['A'], p['W'], nlin)
        gw = gw.reshape(N, -1)
        ds = pg.Dataset(p['S'], p['dt'], p['ibasis'], x_dtype="f32")
        for seg in (100000, 16, 8, 4, 2, 1):
            os.environ["PYGLM_GEMM_SEG"] = str(seg)
            l, b, w = ds.ll_grad(bias, p['w'], p['A'], p['W'], nlin=nlin, path="tc")
            print("N=%d B=%d nlin=%d seg=%6d  ll max rel %.2e  (abs/typ %.2e)  g_bias %.2e  g_w %.2e" % (
                N, B, nlin, seg, np.max(np.abs(l - ll) / np.abs(ll)), np.max(np.abs(l - ll)) / np.median(np.abs(ll)),
                rel_err(b, gb), rel_err(w, gw)), flush=True)
        ds.close()
