"""Accuracy of the tensor-core path vs the FP64 path on the device at C2 size, per flush interval."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theano_pyglm_b200 as pg
from bench import make_inputs, WORKLOADS
wl = WORKLOADS["c2"]; inp = make_inputs(wl, 1234)
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"])
ref = ds.ll_grad(inp["bias"], inp["w"], path="fp64")
for F in sys.argv[1:]:
    os.environ["PYGLM_TC_FLUSH"] = F
    out = ds.ll_grad(inp["bias"], inp["w"], path="tc")
    e = [float(np.max(np.abs(o - r)) / np.max(np.abs(r))) for o, r in zip(out, ref)]
    print("flush", F, "rel err ll %.2e gb %.2e gw %.2e" % tuple(e), "ll elementwise %.2e" % float(np.max(np.abs(out[0]-ref[0])/np.abs(ref[0]))))
