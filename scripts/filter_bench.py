import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import theano_pyglm_b200 as pg
from bench import make_inputs, WORKLOADS
wl = WORKLOADS["c2"]; inp = make_inputs(wl, 1234)
mode = sys.argv[1] if len(sys.argv) > 1 else "f32"          # f32 (X + planes), planes, f64
ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], x_dtype=mode)
st = torch.cuda.current_stream()
for _ in range(3): ds.refilter(st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10): ds.refilter(st.cuda_stream)
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byts = sum(ds.filter_bytes())
print("filter C2 [" + mode + "]: %.1f us  %.1f GB/s (%.1f%% of 6531.9)" % (ms * 1e3, byts / ms / 1e6, byts / ms / 1e6 / 65.319))
