"""Time the fused ll+grad kernel at C2 under the experiment switches given in the environment (PYGLM_TC_*), and dump
its pipeline trace when PYGLM_TC_TRACE is set.  python scripts/fused_probe.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import theano_pyglm_b200 as pg          # noqa: E402
from bench import WORKLOADS, make_inputs  # noqa: E402

wl = WORKLOADS[os.environ.get("PROBE_WORKLOAD", "c2")]
inp = make_inputs(wl, 1234)
N, NB = wl["N"], wl["N"] * wl["B"]
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"])
dev = torch.device("cuda", 0)
d_bias = torch.from_numpy(inp["bias"]).to(dev)
d_w = torch.from_numpy(inp["w"]).to(dev)
d_out = torch.zeros(N * (2 + NB), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream()


def step():
    ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, "explinear", 0, N, "auto", d_out[:N].data_ptr(),
                   d_out[N:2 * N].data_ptr(), d_out[2 * N:].data_ptr(), st.cuda_stream)


for _ in range(8):
    step()
torch.cuda.synchronize()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
best = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / steps)
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("PYGLM_TC"))
print("fused_probe [%s] ms/eval: min %.4f median %.4f  ll0 %.6f" % (tag, min(best), float(np.median(best)), float(d_out[0])))
