"""Categorical draw from unnormalised log-probabilities (interface of pyglm/inference/log_sum_exp.py)."""
import numpy as np


def log_sum_exp_sample(lnp, u=None):
    """Index of the first entry whose cumulative probability reaches u (log_sum_exp.py:26-32).
    `u` defaults to one np.random.rand() draw, exactly what the reference consumes; tests inject it."""
    lnp = np.ravel(np.asarray(lnp, dtype=np.float64))
    assert lnp.ndim == 1, "ERROR: logSumExpSample requires a 1-d vector"
    max_lnp = np.max(lnp)
    with np.errstate(invalid='ignore', divide='ignore'):
        denom = np.log(np.sum(np.exp(lnp - max_lnp))) + max_lnp
        p_safe = np.exp(lnp - denom)
    total = np.sum(p_safe)
    if total == 0 or not np.isfinite(total):
        raise Exception("Invalid input. Probability infinite everywhere.")
    if u is None:
        u = np.random.rand()
    acc = 0.0
    for n in range(lnp.size):
        acc += p_safe[n]
        if u <= acc:
            return n
    raise Exception("Invalid choice in logSumExp!")
