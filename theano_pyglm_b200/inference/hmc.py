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


def hmc_batched(U_and_grad, step_sz, n_steps, q_curr, active=None, adaptive_step_sz=True, tgt_accept_rate=0.9,
                avg_accept_time_const=0.95, avg_accept_rate=None, min_step_sz=1e-5, max_step_sz=1.0, rng=None):
    """M independent HMC chains advanced in lock-step: the same transition as `hmc` for every row of q_curr
    (M, D), with one call of `U_and_grad(Q) -> (U (M,), grad (M, D))` per leapfrog step instead of M.

    The GLM parameters of different neurons are conditionally independent given the network and the data
    (gibbs.py:53-61), and one engine call evaluates the log posterior and gradient of all of them, so the
    per-neuron loop of the reference's HMC updates collapses into n_steps + 2 engine calls.
    step_sz and avg_accept_rate are per-chain arrays (M,); `active` (M, D) masks coordinates that do not move
    (chains with an all-zero row are left alone and keep their step size).  Returns (q_next, step_sz, rate)."""
    rng = np.random if rng is None else rng
    q0 = np.array(q_curr, dtype=np.float64, copy=True)
    M, D = q0.shape
    act = np.ones((M, D)) if active is None else np.asarray(active, dtype=np.float64)
    eps = np.asarray(step_sz, dtype=np.float64).reshape(M, 1)
    rate0 = np.full(M, 0.9) if avg_accept_rate is None else np.asarray(avg_accept_rate, dtype=np.float64)
    U0, g = U_and_grad(q0)
    p0 = rng.randn(M, D) * act
    q = q0.copy()
    p = p0 - 0.5 * eps * g * act
    for i in range(n_steps):
        q = q + eps * p
        U1, g = U_and_grad(q)
        if i < n_steps - 1:
            p = p - eps * g * act
    p = -(p - 0.5 * eps * g * act)
    H0 = U0 + 0.5 * np.sum(p0 ** 2, axis=1)
    H1 = U1 + 0.5 * np.sum(p ** 2, axis=1)
    moving = act.any(axis=1)
    with np.errstate(invalid='ignore'):
        accept = moving & np.isfinite(H1) & (np.log(rng.rand(M)) < H0 - H1)
    q_next = np.where(accept[:, None], q, q0)
    if not adaptive_step_sz:
        return q_next
    rate = np.where(moving, avg_accept_time_const * rate0 + (1.0 - avg_accept_time_const) * accept, rate0)
    step = np.where(moving, eps[:, 0] * np.where(rate > tgt_accept_rate, 1.02, 0.98), eps[:, 0])
    return q_next, np.clip(step, min_step_sz, max_step_sz), rate
