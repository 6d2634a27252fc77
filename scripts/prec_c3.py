"""Tensor-core GEMM path vs the FP64 path on the device at C3 size (N=256, T=1e6, B=5)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theano_pyglm_b200 as pg
from bench import make_inputs, WORKLOADS
wl = dict(WORKLOADS["c3"]);
if len(sys.argv) > 1: wl["T"] = int(sys.argv[1])
inp = make_inputs(wl, 1234)
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"])
t0 = time.time(); ref = ds.ll_grad(inp["bias"], inp["w"], path="fp64"); t1 = time.time()
out = ds.ll_grad(inp["bias"], inp["w"], path="tc"); t2 = time.time()
out = ds.ll_grad(inp["bias"], inp["w"], path="tc"); t3 = time.time()
e = [float(np.max(np.abs(o - r)) / np.max(np.abs(r))) for o, r in zip(out, ref)]
print("C3 T=%d: rel err ll %.2e gb %.2e gw %.2e | ll elementwise %.2e | host-call s: fp64 %.3f tc(first) %.3f tc %.3f"
      % (wl["T"], e[0], e[1], e[2], float(np.max(np.abs(out[0]-ref[0])/np.abs(ref[0]))), t1-t0, t2-t1, t3-t2))
