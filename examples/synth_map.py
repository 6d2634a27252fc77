"""MAP fit of a standard GLM to synthetic data: the call stack of the reference's `python -m test.synth_map`
(test/synth_map.py) on the B200 engine.

    python examples/synth_map.py [N] [T_seconds]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theano_pyglm_b200.inference.coord_descent import coord_descent
from theano_pyglm_b200.utils.synth import make_synth_dataset

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T_stop = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0

model, popn, x_true, data = make_synth_dataset('standard_glm', N=N, T_stop=T_stop, seed=0)
print("simulated %d spikes from %d neurons in %.0f s" % (data['S'].sum(), N, T_stop))
popn.add_data(data)                                   # uploads the spikes, runs the spike-history filter once
lp_true = popn.compute_log_p(x_true)
np.random.seed(1)
x0 = popn.sample()
lp_start = popn.compute_log_p(x0)
t0 = time.perf_counter()
x_map = coord_descent(popn, x0=x0, maxiter=1)         # all neurons fitted together, one engine call per evaluation (x0 is updated in place)
dt = time.perf_counter() - t0
print("log p: start %.1f, MAP %.1f, true parameters %.1f   (%.2f s)" % (lp_start, popn.compute_log_p(x_map), lp_true, dt))
for n in range(min(N, 4)):
    print("neuron %d bias: true %.3f  MAP %.3f" % (n, x_true['glms'][n]['bias']['bias'][0], x_map['glms'][n]['bias']['bias'][0]))
