"""Hamiltonian Monte Carlo with step-size adaptation.

Stands in for `hips.inference.hmc.hmc` (imported at inference/gibbs.py:15, not vendored); same
call shape as the reference's call sites (gibbs.py:305, :748):
    q, new_step_sz, new_accept_rate = hmc(U, grad_U, step_sz, n_steps, q_curr,
                                          adaptive_step_sz=True, avg_accept_rate=rate)
U is the negative log density.
"""
import numpy as np


def hmc(U, grad_U, step_sz, n_steps, q_curr, adaptive_step_sz=False, tgt_accept_rate=0.9,
        avg_accept_time_const=0.95, avg_accept_rate=0.9, min_step_sz=1e-5, max_step_sz=1.0, rng=None):
    rng = np.random if rng is None else rng
    q = np.array(q_curr, dtype=np.float64, copy=True)
    p = rng.randn(*q.shape)
    p_curr = p.copy()
    p = p - 0.5 * step_sz * grad_U(q)                  # leapfrog
    for i in range(n_steps):
        q = q + step_sz * p
        if i < n_steps - 1:
            p = p - step_sz * grad_U(q)
    p = -(p - 0.5 * step_sz * grad_U(q))
    H_curr = U(q_curr) + 0.5 * np.sum(p_curr ** 2)
    H_prop = U(q) + 0.5 * np.sum(p ** 2)
    accept = np.isfinite(H_prop) and np.log(rng.rand()) < H_curr - H_prop
    q_next = q if accept else np.array(q_curr, dtype=np.float64, copy=True)
    if not adaptive_step_sz:
        return q_next
    rate = avg_accept_time_const * avg_accept_rate + (1.0 - avg_accept_time_const) * float(accept)
    step = step_sz * (1.02 if rate > tgt_accept_rate else 0.98)
    return q_next, float(np.clip(step, min_step_sz, max_step_sz)), rate
