"""MCMC over a sparse weighted network GLM: the call stack of the reference's `test.synth_mcmc`
 on the B200 engine -- lock-step HMC on the GLM parameters, collapsed Gibbs over A / W.

    python examples/synth_mcmc.py [N] [T_seconds] [samples]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theano_pyglm_b200.inference.gibbs import gibbs_sample
from theano_pyglm_b200.utils.synth import make_synth_dataset

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T_stop = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
n_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 10

model, popn, x_true, data = make_synth_dataset('sparse_weighted_model', N=N, T_stop=T_stop, seed=0)
print("simulated %d spikes; true network has %d edges" % (data['S'].sum(), x_true['net']['graph']['A'].sum()))
popn.add_data(data)
t0 = time.perf_counter()
samples = gibbs_sample(popn, N_samples=n_samples, init_from_mle=True)
print("%d samples in %.1f s" % (n_samples, time.perf_counter() - t0))
A_mean = np.mean([s['net']['graph']['A'] for s in samples[n_samples // 2:]], axis=0)
A_true = x_true['net']['graph']['A']
print("posterior edge probability: %.2f on true edges, %.2f elsewhere" %
      (A_mean[A_true == 1].mean(), A_mean[A_true == 0].mean() if (A_true == 0).any() else float('nan')))
print("log p of the last sample %.1f, of the true parameters %.1f" % (popn.compute_log_p(samples[-1]), popn.compute_log_p(x_true)))
