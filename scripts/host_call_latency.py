"""Latency of the host-buffer entry point pyglm_b200_ll_grad (what Population.ll_grad calls) at C2."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theano_pyglm_b200 as pg
from bench import WORKLOADS, make_inputs

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
inp = make_inputs(wl, 1234)
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"])
for grad in (True, False):
    for _ in range(5):
        ds.ll_grad(inp["bias"], inp["w"], grad=grad)
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        out = ds.ll_grad(inp["bias"], inp["w"], grad=grad)
    dt = (time.perf_counter() - t0) / n
    print("host entry, grad=%s: %.1f us per call (%.0f evals/s)" % (grad, dt * 1e6, 1.0 / dt), flush=True)
