"""Synthetic data sets (interface of test/generate_synth_data.py:56-135)."""
import numpy as np

from ..models.model_factory import check_stability, make_model, stabilize_sparsity
from ..population import Population


def gen_synth_data(N, T_stop, popn, x_true, dt=0.001, dt_stim=0.1, D_stim=1, stim=None):
    """Simulate and package the data dict (generate_synth_data.py:56-79)."""
    S, X = popn.simulate(x_true, (0, T_stop), dt, stim, dt_stim)
    return {"S": S, "X": X, "N": N, "dt": dt, "T": float(T_stop), "stim": stim, 'dt_stim': dt_stim,
            'vars': x_true}


def make_synth_dataset(model_name, N, T_stop, dt=0.001, seed=None, device=0, max_tries=20):
    """run_gen_synth_data (generate_synth_data.py:81-135) without the file I/O: model, population, true
    parameters from the prior (re-drawn until the network is stable) and a simulated recording."""
    if seed is not None:
        np.random.seed(seed)
    model = make_model(model_name, N=N, dt=dt)
    stabilize_sparsity(model)
    popn = Population(model, device=device)
    for _ in range(max_tries):
        x_true = popn.sample()
        if check_stability(model, x_true, N):
            break
    else:
        raise RuntimeError("Sampled network is unstable!")
    dt_stim = 0.1
    D = model['bkgd'].get('D_stim', 1) if model['bkgd']['type'].lower() == 'basis' else 1
    stim = np.random.randn(int(T_stop / dt_stim), D)
    data = gen_synth_data(N, T_stop, popn, x_true, dt, dt_stim, D, stim)
    return model, popn, x_true, data
