"""Where does the GEMM forward kernel's time go?  Times ll-only evaluations (forward kernel + tiny reductions) at C3
under the PYGLM_GEMM_DEBUG knobs: 1 = no MMAs, 2 = every tile loads the same (L2-hot) rows, 4 = no epilogue math."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theano_pyglm_b200 as pg
from bench import WORKLOADS, make_inputs

wl = WORKLOADS["c3"]
inp = make_inputs(wl, 1234)
N = wl["N"]
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"])
dev = torch.device("cuda", 0)
d_bias = torch.from_numpy(inp["bias"]).to(dev)
d_w = torch.from_numpy(inp["w"]).to(dev)
d_ll = torch.zeros(N, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream()
for mode in (0, 3):
    for dbg in (0, 1):
        os.environ["PYGLM_GEMM_MODE"] = str(mode)
        os.environ["PYGLM_GEMM_DEBUG"] = str(dbg)
        for _ in range(3):
            ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, "explinear", 0, N, "tc", d_ll.data_ptr(), 0, 0, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, "explinear", 0, N, "tc", d_ll.data_ptr(), 0, 0, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        print("mode %d debug %d: %.3f ms per forward" % (mode, dbg, e0.elapsed_time(e1) / 10), flush=True)
