"""Where does one full MCMC iteration (gibbs_sample, batched schedule) spend its time?  N=64, T=2e5."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theano_pyglm_b200.inference.gibbs import gibbs_sample
from theano_pyglm_b200.models.model_factory import make_model, stabilize_sparsity
from theano_pyglm_b200.population import Population

N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 200_000
model = make_model('sparse_weighted_model', N=N, dt=0.001)
stabilize_sparsity(model)
popn = Population(model)
rng = np.random.default_rng(0)
S = (rng.random((T, N)) < 0.02).astype(float)
popn.add_data({'S': S, 'N': N, 'dt': 0.001, 'T': T * 0.001, 'stim': None, 'dt_stim': 0.1})
np.random.seed(1)
x0 = popn.sample()
gibbs_sample(popn, N_samples=1, x0=x0)            # warm-up (allocations, graph capture)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
gibbs_sample(popn, N_samples=2, x0=x0)
pr.disable()
print("2 iterations: %.2f s" % (time.perf_counter() - t0))
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
