"""MAP estimation by coordinate descent (interface of pyglm/inference/coord_descent.py).

`coord_descent(population, x0, maxiter, atol)` alternates fitting the GLMs and the network exactly
like the reference (:206-266).  Two GLM fitters:
  * `fit_glm`          one neuron, scipy BFGS with maxiter 225 and the reference's NaN guards
                       (:161-204); every function / gradient evaluation is one engine call.
  * `fit_glms_batched` all N neurons in lock-step (default): the N problems are independent given the
                       network (population.py:28-32), and one engine call evaluates ll + gradient for
                       every neuron at the cost of one, so a limited-memory BFGS runs on all of them at
                       once -- one kernel launch per line-search trial instead of N.
"""
import copy

import numpy as np
import scipy.optimize as opt

from .smart_init import initialize_with_data


def fit_glm(population, x, n, maxiter=225, disp=False):
    """Fit neuron n's GLM parameters in place (coord_descent.py:161-204)."""
    xn = x['glms'][n]

    def both(vec):
        population.set_glm_param_vector(xn, vec)
        return population.glm_log_p_grad(x, n)

    def nll(vec):
        y = -both(vec)[0]
        return 1e16 if np.isnan(y) else y              # :170-173

    def grad_nll(vec):
        g = -both(vec)[1]
        return np.zeros_like(g) if np.any(np.isnan(g)) else g     # :176-180

    res = opt.minimize(nll, population.glm_param_vector(xn), method="bfgs", jac=grad_nll,
                       options={'disp': disp, 'maxiter': maxiter})
    population.set_glm_param_vector(xn, res.x)
    return res


def fit_glms_batched(population, x, maxiter=225, gtol=1e-5, history=20, verbose=False, n_lo=0, n_hi=None):
    """Lock-step L-BFGS over the neurons [n_lo, n_hi) (all by default; a neuron-sharded rank passes its own block);
    returns the number of iterations taken."""
    n_hi = population.N if n_hi is None else n_hi
    N = n_hi - n_lo
    P_all = np.stack([population.glm_param_vector(x['glms'][n]) for n in range(population.N)])

    def evaluate(P):
        P_all[n_lo:n_hi] = P
        lp, g = population.glms_log_p_grad_dense(P_all, x, n_lo, n_hi)   # priors included, no per-neuron dict traffic
        f = np.where(np.isnan(lp), 1e16, -lp)              # same guards as fit_glm
        g = -g
        g[np.any(np.isnan(g), axis=1)] = 0.0
        return f, g

    P = P_all[n_lo:n_hi].copy()
    f, g = evaluate(P)
    S_hist, Y_hist = [], []                                 # lists of (N, D) arrays
    active = np.ones(N, dtype=bool)
    it = 0
    for it in range(1, maxiter + 1):
        active &= np.max(np.abs(g), axis=1) > gtol
        if not active.any():
            break
        # two-loop recursion, per neuron (rows are independent problems)
        q = g.copy()
        alphas = []
        for s_, y_ in zip(reversed(S_hist), reversed(Y_hist)):
            rho = 1.0 / np.maximum(np.sum(s_ * y_, axis=1), 1e-300)
            a = rho * np.sum(s_ * q, axis=1)
            q -= a[:, None] * y_
            alphas.append((a, rho))
        if S_hist:
            gamma = np.sum(S_hist[-1] * Y_hist[-1], axis=1) / np.maximum(np.sum(Y_hist[-1] ** 2, axis=1), 1e-300)
            q *= gamma[:, None]
        else:
            q *= (1.0 / np.maximum(np.linalg.norm(g, axis=1), 1.0))[:, None]
        for (a, rho), s_, y_ in zip(reversed(alphas), S_hist, Y_hist):
            b = rho * np.sum(y_ * q, axis=1)
            q += (a - b)[:, None] * s_
        d = -q
        gd = np.sum(g * d, axis=1)
        bad = gd >= 0                                       # not a descent direction: fall back to steepest descent
        d[bad] = -g[bad]
        gd[bad] = -np.sum(g[bad] ** 2, axis=1)
        d[~active] = 0.0
        # backtracking (Armijo) line search, all neurons at once
        step = np.ones(N)
        done = ~active
        f_new, g_new, P_new = f.copy(), g.copy(), P.copy()
        for _ in range(30):
            trial = P + np.where(done, 0.0, step)[:, None] * d
            ft, gt = evaluate(trial)
            ok = (~done) & (ft <= f + 1e-4 * step * gd)
            f_new[ok], g_new[ok], P_new[ok] = ft[ok], gt[ok], trial[ok]
            done |= ok
            if done.all():
                break
            step[~done] *= 0.5
        stalled = ~done                                     # no acceptable step: freeze that neuron
        active &= ~stalled
        s_ = P_new - P
        y_ = g_new - g
        curv = np.sum(s_ * y_, axis=1) > 1e-12              # keep the pair only where the curvature is positive
        s_[~curv] = 0.0
        y_[~curv] = 0.0
        S_hist.append(s_); Y_hist.append(y_)
        if len(S_hist) > history:
            S_hist.pop(0); Y_hist.pop(0)
        P, f, g = P_new, f_new, g_new
        if verbose:
            print("L-BFGS iter %d: sum LP %.3f, active %d" % (it, -f.sum(), int(active.sum())))
    for n in range(N):
        population.set_glm_param_vector(x['glms'][n_lo + n], P[n])
    return it


def fit_network(population, x):
    """Fit the differentiable network variables against the network prior (coord_descent.py:84-159: Newton-CG on
    `-network.log_prior`; the data never enter).  Constant weights: nothing to fit (:141).  Gaussian weights: the
    maximiser of the Gaussian prior is its mean, written here in closed form -- off-diagonal `prior.mu`, diagonal
    `refractory_prior.mu` (weights.py:47-71).  (The reference itself stops with an AttributeError on this branch:
    `Network` has `log_p`, not the `log_prior` that coord_descent.py:113 evaluates; SURVEY.md appendix A.)"""
    wm = population.network.weights
    if 'W' not in x['net']['weights'] or not hasattr(wm, 'prior'):
        return x
    N = population.N
    W = np.full((N, N), float(wm.prior.mu.get_value()))
    if hasattr(wm, 'refractory_prior'):
        W[np.diag_indices(N)] = float(wm.refractory_prior.mu.get_value())
    x['net']['weights']['W'] = W.ravel()
    return x


def coord_descent(population, x0=None, maxiter=50, atol=1e-5, batched=True, verbose=False):
    """MAP estimate by coordinate descent (coord_descent.py:206-266)."""
    N = population.model['N']
    if x0 is None:
        x0 = population.sample()
    initialize_with_data(population, population.data_sequences[-1], x0)
    x = x0
    lp_prev = population.compute_log_p(x)
    if verbose:
        print("Initial LP=%.2f." % lp_prev)
    converged, it = False, 0
    while not converged and it < maxiter:
        it += 1
        if batched:
            fit_glms_batched(population, x, verbose=verbose)
        else:
            for n in range(N):
                fit_glm(population, x, n)
        fit_network(population, x)
        lp = population.compute_log_p(x)
        if verbose:
            print("Iteration %d: LP=%.2f. Change in LP: %.2f" % (it, lp, lp - lp_prev))
        converged = np.abs(lp - lp_prev) < atol
        lp_prev = lp
    return x
