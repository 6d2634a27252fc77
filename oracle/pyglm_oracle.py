"""CPU oracle for the pyglm hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy float64 restatement of the reference algorithm (slinderman/theano_pyglm,
mounted read-only at /root/reference) for the one hot path this repo accelerates:
spike-history filtering -> activation -> nonlinearity -> Poisson log-likelihood ->
gradient, plus the collapsed spike-and-slab Gibbs update over one column of A/W.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (theano_pyglm_b200) never does.

PARITY UNPINNED against Theano itself: the reference is Python-2 + Theano-0.6 and
cannot be imported or executed in this image, and it ships
no golden vectors.  What *is* pinned here:
  * the filter calls scipy.signal.fftconvolve exactly as utils/basis.py:232 does and
    is cross-checked against a direct causal sum (tests/test_oracle.py);
  * the reference's only numeric assertion on this path, lam(graph) == f_nlin(X_sim)
    (test/generate_synth_data.py:124-129), is reproduced by `simulate` +
    `population_activation`;
  * hand-derived gradients are checked against torch.autograd (float64) and central
    finite differences of `glm_ll`, i.e. against the definition T.grad implements;
  * Gauss-Hermite constants / logsumexp come from the same numpy/scipy calls.

Every function cites the reference file:line it follows (paths relative to
/root/reference/).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.signal
from scipy.special import logsumexp

NLIN_EXP = 0
NLIN_SOFTPLUS = 1  # the reference calls this 'explinear' (components/nlin.py:42-47)


# --------------------------------------------------------------------------------------
# Basis construction  (pyglm/utils/basis.py:56-106, components/impulse.py:92-112,359-376)
# --------------------------------------------------------------------------------------
def create_cosine_basis(prms):
    """Raised-cosine basis on 100 points.  pyglm/utils/basis.py:56-106."""
    n_pts = 100
    n_cos = prms['n_cos']
    n_eye = prms['n_eye']
    n_bas = n_eye + n_cos
    basis = np.zeros((n_pts, n_bas))
    basis[:n_eye, :n_eye] = np.eye(n_eye)                                  # :74

    a = prms['a']
    b = prms['b']
    u_ir = np.log(a * np.arange(n_pts) + b)                                # :81-82
    ctrs = u_ir[np.floor(np.linspace(n_eye, (n_pts / 2.0), n_cos)).astype(int)]   # :83
    if len(ctrs) == 1:
        w = ctrs / 2
    else:
        w = (ctrs[-1] - ctrs[0]) / (n_cos - 1)                             # :87

    def basis_fn(u, c, w):                                                 # :90
        return (np.cos(np.maximum(-np.pi, np.minimum(np.pi, (u - c) * np.pi / w / 2.0))) + 1) / 2.0

    for i in range(n_cos):
        basis[:, n_eye + i] = basis_fn(u_ir, ctrs[i], w)

    if prms['orth']:
        basis = scipy.linalg.orth(basis)                                   # :97-98
    if prms['norm']:
        if np.any(basis < 0):
            raise Exception("We can only normalize nonnegative impulse responses!")
        basis = basis / np.tile(np.sum(basis, axis=0), [n_pts, 1]) / (1.0 / n_pts)   # :104
    return basis


def create_basis(prms):
    """pyglm/utils/basis.py:9-26 (cosine / identity only; the configs use cosine)."""
    typ = prms['type'].lower()
    if typ == 'cosine':
        return create_cosine_basis(prms)
    if typ in ('identity', 'eye'):
        return np.eye(prms['n_eye'])                                       # :188-199
    raise Exception("Unrecognized basis type: %s" % typ)


def interpolate_basis_linear(basis, dt, dt_max, norm):
    """LinearBasisImpulses.initialize_basis, components/impulse.py:92-112."""
    (L, B) = basis.shape
    Lt_int = int(dt_max / dt)                                              # :99
    t_int = np.linspace(0, 1, Lt_int)
    t_bas = np.linspace(0, 1, L)
    ibasis = np.zeros((len(t_int), B))
    for b in range(B):
        ibasis[:, b] = np.interp(t_int, t_bas, basis[:, b])
    if norm:
        ibasis = ibasis / dt_max                                           # :107-108
    return ibasis


def interpolate_basis_dirichlet(basis, dt, dt_max, norm):
    """DirichletImpulses.initialize_basis, components/impulse.py:359-376."""
    (L, B) = basis.shape
    t_int = np.arange(0.0, dt_max, step=dt)                                # :365
    t_bas = np.linspace(0.0, dt_max, L)                                    # :367
    ibasis = np.zeros((len(t_int), B))
    for b in range(B):
        ibasis[:, b] = np.interp(t_int, t_bas, basis[:, b])
    if norm:
        ibasis = ibasis / np.trapezoid(ibasis, t_int, axis=0)              # :373-374 (np.trapz)
    return ibasis


# --------------------------------------------------------------------------------------
# K1: spike-history filtering  (pyglm/utils/basis.py:201-236)
# --------------------------------------------------------------------------------------
def convolve_with_basis(stim, basis):
    """fS[t,d,b] = sum_{k=1..R} basis[k-1,b] * stim[t-k,d].  utils/basis.py:201-236.

    Same call shape as the reference: one zero row prepended (:220) and one
    scipy.signal.fftconvolve(..., 'full')[:T] per basis column (:232-234)."""
    (T, D) = stim.shape
    (R, B) = basis.shape
    basis = np.vstack((np.zeros((1, B)), basis))
    fstim = np.empty((T, D, B))
    for b in range(B):
        fstim[:, :, b] = scipy.signal.fftconvolve(
            stim, np.reshape(basis[:, b], [R + 1, 1]), 'full')[:T, :]
    return fstim


def convolve_with_basis_direct(stim, basis):
    """Same quantity by the defining causal sum, exploiting sparse `stim`; float64.
    Summation order: increasing spike time (== decreasing lag), the order the CUDA
    filter kernel also uses, so float64 sums agree bit-for-bit with it."""
    stim = np.asarray(stim)
    (T, D) = stim.shape
    (R, B) = basis.shape
    out = np.zeros((T, D, B))
    ts, ds = np.nonzero(stim)
    order = np.lexsort((ts, ds))        # by column, then increasing time
    ts, ds = ts[order], ds[order]
    cnt = stim[ts, ds].astype(np.float64)
    for t0, d, c in zip(ts, ds, cnt):
        hi = min(T, t0 + 1 + R)
        n = hi - (t0 + 1)
        if n > 0:
            out[t0 + 1:hi, d, :] += c * basis[:n, :]
    return out


# --------------------------------------------------------------------------------------
# Components: nonlinearity, impulse currents, network weights
# --------------------------------------------------------------------------------------
def nlin(x, nlin_type):
    """components/nlin.py:25 (exp) and :43,47 (log(1+exp(x)), named 'explinear')."""
    return nlin_and_derivative(x, nlin_type)[0]


def _softplus_parts(x):
    """lam = log(1+e^x), f' = sigmoid(x), log lam and f'/lam, each formed stably.  For x < -30, lam = e q with
    q = log1p(e)/e = 1 - e/2 + e^2/3 - ..., so log lam = x + log q and f'/lam = 1/((1+e) q) stay finite however negative
    x is.  (The reference's literal log(log(1+exp(x))) is -inf below x ~ -37 and its ll NaN there; its callers map
    that NaN to -inf, gibbs.py:1011-1012.  The finite limit kept here is > 700 below any competing term.)"""
    e = np.exp(-np.abs(x))
    l1p = np.log1p(e)
    pos = x > 0
    lam = np.where(pos, x + l1p, l1p)
    sig = np.where(pos, 1.0, e) / (1.0 + e)
    with np.errstate(divide='ignore', invalid='ignore'):
        loglam = np.log(lam)
        ratio = sig / lam
    far = x < -30.0                                   # rare: patch only those entries
    if np.any(far):
        ef = e[far]
        q = 1.0 - ef * (0.5 - ef / 3.0)
        loglam[far] = x[far] + np.log(q)
        ratio[far] = 1.0 / ((1.0 + ef) * q)
    return lam, sig, loglam, ratio


def nlin_and_derivative(x, nlin_type):
    """Returns (lam, dlam/dx, log lam)."""
    x = np.asarray(x, dtype=np.float64)
    if nlin_type == NLIN_EXP:
        lam = np.exp(x)
        return lam, lam, x.copy()
    return _softplus_parts(x)[:3]


def poisson_residual(x, S_n, dt, nlin_type):
    """r_t = d/dx_t of (-dt*lam + S log lam) = S * f'/lam - dt * f'  (what T.grad propagates through glm.py:52), with
    f'/lam formed without dividing by an underflowed lam (exp: f'/lam = 1)."""
    x = np.asarray(x, dtype=np.float64)
    if nlin_type == NLIN_EXP:
        return S_n - dt * np.exp(x)
    _lam, sig, _ll, ratio = _softplus_parts(x)
    return S_n * ratio - dt * sig


def _ll_and_residual(x, S_n, dt, nlin_type):
    """(per-bin ll terms summed over t, residual r) in one pass over the nonlinearity."""
    x = np.asarray(x, dtype=np.float64)
    if nlin_type == NLIN_EXP:
        lam = np.exp(x)
        return np.sum(-dt * lam + x * S_n, axis=0), S_n - dt * lam
    lam, sig, loglam, ratio = _softplus_parts(x)
    return np.sum(-dt * lam + loglam * S_n, axis=0), S_n * ratio - dt * sig


def dirichlet_beta(g):
    """beta = |g| / sum|g| per presynaptic neuron.  components/impulse.py:286-291.
    g: (N_pre, B) -> (N_pre, B)."""
    gabs = np.abs(g)
    return gabs / np.sum(gabs, axis=1, keepdims=True)


def impulse_current(fS, w):
    """I_imp[t,pre] = sum_b ir[t,pre,b] * w[pre,b].  components/impulse.py:45-58 (:308)."""
    return np.sum(fS * w[None, :, :], axis=2)


def effective_weights(A, W, n):
    """W_eff = A[:,n] * W[:,n].  glm.py:31-35."""
    return A[:, n].astype(np.float64) * W[:, n]


# --------------------------------------------------------------------------------------
# K2: log-likelihood and gradient  (glm.py:31-52; gradient == T.grad of that, coord_descent.py:22-30)
# --------------------------------------------------------------------------------------
def glm_ll_from_activation(x, S_n, dt, nlin_type):
    """ll = sum_t(-dt*lam + log(lam)*S).  glm.py:43-52."""
    lam, _, loglam = nlin_and_derivative(x, nlin_type)
    with np.errstate(invalid='ignore'):
        return float(np.sum(-dt * lam + loglam * S_n))      # NaN/inf propagate like Theano's


def glm_ll(fS, S, dt, n, bias_n, w_n, A, W, nlin_type, I_stim=0.0):
    """Reference-shaped single-neuron evaluation (materialises T x N x B like impulse.py:58).
    w_n: (N_pre, B) impulse weights of neuron n (already beta-normalised for Dirichlet)."""
    I_imp = impulse_current(fS, w_n)                      # impulse.py:58
    I_net = I_imp @ effective_weights(A, W, n)            # glm.py:39
    x = bias_n + I_stim + I_net                           # glm.py:43-45
    return glm_ll_from_activation(x, S[:, n], dt, nlin_type)


def glm_ll_grad(fS, S, dt, n, bias_n, w_n, A, W, nlin_type, I_stim=0.0):
    """ll and d ll / d(bias, w_ir) for neuron n; gradient order = sorted keys
    'bias' < 'imp' (theano_func_wrapper.py:53-67, packvec.py:23).
    r_t = (S_t/lam_t - dt) * f'(x_t);  d/dbias = sum r;  d/dw[pre,b] = W_eff[pre] * sum_t ir[t,pre,b] r_t."""
    T, N, B = fS.shape
    weff = effective_weights(A, W, n)
    I_imp = impulse_current(fS, w_n)
    x = bias_n + I_stim + I_imp @ weff
    Sn = S[:, n].astype(np.float64)
    ll, r = _ll_and_residual(x, Sn, dt, nlin_type)
    ll = float(ll)
    g_bias = float(np.sum(r))
    G = np.tensordot(r, fS, axes=(0, 0))                  # (N_pre, B)
    g_w = weff[:, None] * G
    return ll, g_bias, g_w


def population_activation(fS, bias, w, A, W):
    """Whole-population GEMM form: x[t,n] = bias[n] + X[t,:] @ M[:,n],
    M[(pre,b),n] = A[pre,n] W[pre,n] w[n,pre,b].   Same algebra as glm.py:33-45."""
    T, N, B = fS.shape
    Weff = A.astype(np.float64) * W                       # (pre, post)
    M = (np.transpose(w, (1, 2, 0)) * Weff[:, None, :]).reshape(N * B, N)
    return bias[None, :] + fS.reshape(T, N * B) @ M


def population_ll_grad(fS, S, dt, bias, w, A, W, nlin_type, fstim=None, w_stim=None):
    """Best-effort CPU form (multithreaded BLAS): two GEMMs for all N neurons.
    bias (N,), w (N_post, N_pre, B).  Returns ll (N,), g_bias (N,), g_w (N_post,N_pre,B).
    With a filtered stimulus fstim (T, F) and weights w_stim (N, F) the activation gains
    I_stim = fstim @ w_stim[n] (bkgd.py:81) and a fourth output g_w_stim (N, F) = (fstim^T r)^T."""
    T, N, B = fS.shape
    X = fS.reshape(T, N * B)
    x = population_activation(fS, bias, w, A, W)
    if fstim is not None:
        x = x + fstim @ w_stim.T
    Sf = S.astype(np.float64)
    ll, r = _ll_and_residual(x, Sf, dt, nlin_type)
    g_bias = np.sum(r, axis=0)
    G = (X.T @ r).reshape(N, B, N)                        # [(pre,b), post]
    Weff = A.astype(np.float64) * W
    g_w = np.transpose(G, (2, 0, 1)) * Weff.T[:, :, None]
    if fstim is not None:
        return ll, g_bias, g_w, (fstim.T @ r).T
    return ll, g_bias, g_w


# --------------------------------------------------------------------------------------
# Stimulus features  (components/bkgd.py:45-157, BasisStimulus)
# --------------------------------------------------------------------------------------
def interpolate_stim_basis(basis, dt, dt_max, norm):
    """bkgd.py:102-121: resample the basis on np.linspace(0,1,dt_max/dt); if norm, each column is
    divided by its SUM (not trapz, unlike impulse.py:372)."""
    L, B = basis.shape
    Lt_int = int(round(dt_max / dt))
    t_int = np.linspace(0, 1, Lt_int)
    t_bas = np.linspace(0, 1, L)
    ibasis = np.zeros((Lt_int, B))
    for b in range(B):
        ibasis[:, b] = np.interp(t_int, t_bas, basis[:, b])
    if norm:
        ibasis = ibasis / np.sum(ibasis, axis=0)[None, :]
    return ibasis


def filter_stimulus(stim, dt_stim, nT, dt, ibasis):
    """bkgd.py:134-154: interpolate the stimulus onto the spike bins, project it on the basis
    (utils/basis.py:201-236) and flatten to fstim[t, d*B+b]."""
    D = stim.shape[1]
    t = dt * np.arange(nT)
    t_stim = dt_stim * np.arange(stim.shape[0])
    istim = np.zeros((nT, D))
    for d in range(D):
        istim[:, d] = np.interp(t, t_stim, stim[:, d])
    cstim = convolve_with_basis(istim, ibasis)             # (nT, D, B)
    return istim, cstim.reshape(nT, D * cstim.shape[2])


def stim_log_prior(w_stim):
    """bkgd.py:76: spherical Gaussian with the hard-coded sigma 0.01."""
    return np.sum(-0.5 / (0.01 ** 2) * (w_stim - 0.0) ** 2)


def dirichlet_chain_rule(g, g_beta):
    """d ll/d g from d ll/d beta for beta=|g|/sum|g| (impulse.py:286-291).
    g, g_beta: (..., B)."""
    s = np.sum(np.abs(g), axis=-1, keepdims=True)
    beta = np.abs(g) / s
    inner = np.sum(g_beta * beta, axis=-1, keepdims=True)
    return np.sign(g) * (g_beta - inner) / s


# --------------------------------------------------------------------------------------
# Priors  (bias.py:33, priors.py:139,202, impulse.py:320-322, graph.py:68-71, weights.py:67-71)
# --------------------------------------------------------------------------------------
def bias_log_prior(bias_n, mu, sigma):
    return -0.5 / sigma ** 2 * (bias_n - mu) ** 2                          # bias.py:33


def bias_log_prior_grad(bias_n, mu, sigma):
    return -(bias_n - mu) / sigma ** 2


def gaussian_log_p(value, mu, sigma):
    return -0.5 / sigma ** 2 * np.sum((value - mu) ** 2)                   # priors.py:139


def group_lasso_log_p(w_n, mu, sigma, lam):
    """-lam * sum_pre || (w[pre,:]-mu)/sigma ||_2.  priors.py:202."""
    return -1.0 * lam * np.sum(np.sqrt(np.sum(((w_n - mu) / sigma) ** 2, axis=1)))


def group_lasso_log_p_grad(w_n, mu, sigma, lam):
    """NaN at a zero group, exactly like T.grad of sqrt at 0 (SURVEY 'Hard parts')."""
    z = (w_n - mu) / sigma
    nrm = np.sqrt(np.sum(z ** 2, axis=1, keepdims=True))
    with np.errstate(invalid='ignore', divide='ignore'):
        return -lam * z / nrm / sigma


def dirichlet_impulse_log_p(g, alpha):
    """sum_pre (alpha-1) sum log|g| - sum |g|.  impulse.py:320-322."""
    return float(np.sum((alpha - 1.0) * np.log(np.abs(g)) - np.abs(g)))


def erdos_renyi_log_p(A, rho):
    """graph.py:68-71."""
    return float(np.sum(A * np.log(np.minimum(1.0 - 1e-8, rho)) +
                        (1 - A) * np.log(np.maximum(1e-8, 1.0 - rho))))


def gaussian_weight_log_p(W, mu, sigma, mu_ref=None, sigma_ref=None):
    """weights.py:67-71: separate refractory prior on the diagonal."""
    N = W.shape[0]
    if mu_ref is None:
        return gaussian_log_p(W, mu, sigma)
    diag = np.eye(N, dtype=bool)
    return gaussian_log_p(W[~diag], mu, sigma) + gaussian_log_p(W[diag], mu_ref, sigma_ref)


# --------------------------------------------------------------------------------------
# K4: collapsed Gibbs over one column of A/W  (inference/gibbs.py:775-1250, log_sum_exp.py:4-37)
# --------------------------------------------------------------------------------------
DEG_GAUSS_HERMITE = 10
GAUSS_HERMITE_ABSCISSAE, GAUSS_HERMITE_WEIGHTS = np.polynomial.hermite.hermgauss(DEG_GAUSS_HERMITE)  # gibbs.py:787-789


def gibbs_glm_ll(I_bias, I_stim, I_net_other, u_pre, w, S_n, dt, nlin_type):
    """_glm_ll, gibbs.py:910-937: I_net = I_other + w * I_imp[:,n_pre] (:914), then glm.ll."""
    I_net = I_net_other + w * u_pre
    return glm_ll_from_activation(I_bias + I_stim + I_net, S_n, dt, nlin_type)


def log_sum_exp_sample(lnp, u):
    """log_sum_exp.py:4-37 with the single np.random.rand() (:26) injected as `u`."""
    lnp = np.ravel(np.asarray(lnp, dtype=np.float64))
    max_lnp = np.max(lnp)
    denom = np.log(np.sum(np.exp(lnp - max_lnp))) + max_lnp
    p_safe = np.exp(lnp - denom)
    sum_p_safe = np.sum(p_safe)
    if sum_p_safe == 0 or not np.isfinite(sum_p_safe):
        raise Exception("Invalid input. Probability infinite everywhere.")
    acc = 0.0
    for n in range(lnp.size):
        acc += p_safe[n]
        if u <= acc:
            return n
    raise Exception("Invalid choice in logSumExp!")


def gh_candidates(mu_w, sigma_w):
    """W_nns = sqrt(2)*sigma_w*x_GH + mu_w.  gibbs.py:1004."""
    return np.sqrt(2) * sigma_w * GAUSS_HERMITE_ABSCISSAE + mu_w


def collapsed_edge_log_odds(log_L, ll_noA, p_A):
    """gibbs.py:1002-1035 given the 10 quadrature log-likelihoods and the w=0 one.
    Returns (log_pr_noA, log_pr_A)."""
    log_L = np.array(log_L, dtype=np.float64)
    log_L[np.isnan(log_L)] = -np.inf                                        # :1011-1012
    with np.errstate(divide='ignore'):
        weighted = log_L + np.log(GAUSS_HERMITE_WEIGHTS / np.sqrt(np.pi))  # :1015
    weighted[np.isnan(weighted)] = -np.inf
    log_G = logsumexp(weighted)                                             # :1022
    if not np.isfinite(log_G):
        raise Exception("log_G not finie")                                 # :1023-1025
    with np.errstate(divide='ignore'):
        log_pr_A = np.log(p_A) + log_G                                      # :1028
        log_pr_noA = np.log(1.0 - p_A) + ll_noA                             # :1030-1032
    if np.isnan(log_pr_noA):
        log_pr_noA = -np.inf
    return log_pr_noA, log_pr_A


def collapsed_column_sweep(fS, S, dt, n_post, bias_n, w_n, A, W, p_A, nlin_type,
                           mu_w, sigma_w, mu_w_ref, sigma_w_ref,
                           order, uniforms, w_draw):
    """One `update(x, n)` of CollapsedGibbsNetworkColumnUpdate, gibbs.py:1229-1250, with every
    random draw injected:
      order    : the shuffled presynaptic order (:1236-1237)
      uniforms : one uniform per edge for log_sum_exp_sample (:1039)
      w_draw   : callable(n_pre, A_new, mu_w, sigma_w, W_nns, log_L) -> new W[n_pre,n_post];
                 stands in for ARS when A=1 (:1054, hips is un-vendored -> parity unpinned)
                 and for mu + sigma*randn when A=0 (:1063).
    A (int8 N x N) and W (float64 N x N) are modified in place, like the reference.
    Returns a list of per-edge records (log_L[10], ll_noA, log_pr_noA, log_pr_A, A_new)."""
    I_imp = impulse_current(fS, w_n)                       # _precompute_vars :812-833
    S_n = S[:, n_post].astype(np.float64)
    rec = []
    for i, n_pre in enumerate(order):
        # _precompute_other_current :835-864 : gemv with A[n_pre,n_post] = 0
        a_save = A[n_pre, n_post]
        A[n_pre, n_post] = 0
        I_other = I_imp @ effective_weights(A, W, n_post)
        A[n_pre, n_post] = a_save
        if n_pre == n_post:                                # :984-989
            mu, sig = mu_w_ref, sigma_w_ref
        else:
            mu, sig = mu_w, sigma_w
        W_nns = gh_candidates(mu, sig)
        log_L = np.array([gibbs_glm_ll(bias_n, 0.0, I_other, I_imp[:, n_pre], wq, S_n, dt, nlin_type)
                          for wq in W_nns])
        ll_noA = gibbs_glm_ll(bias_n, 0.0, I_other, I_imp[:, n_pre], 0.0, S_n, dt, nlin_type)
        lp_noA, lp_A = collapsed_edge_log_odds(log_L, ll_noA, p_A[n_pre, n_post])
        a_new = log_sum_exp_sample([lp_noA, lp_A], uniforms[i])            # :1039
        A[n_pre, n_post] = a_new
        W[n_pre, n_post] = w_draw(n_pre, a_new, mu, sig, W_nns, log_L)
        rec.append(dict(n_pre=int(n_pre), log_L=log_L, ll_noA=ll_noA,
                        log_pr_noA=lp_noA, log_pr_A=lp_A, A=int(a_new)))
    return rec


# --------------------------------------------------------------------------------------
# Data generation for config C1  (population.py:233-389; generate_synth_data.py:56-79)
# --------------------------------------------------------------------------------------
def simulate(bias, imps, A, W, nT, dt, nlin_type, rng):
    """Time-rescaling spike generator, population.py:290-364 (no stimulus).
    imps: (N_pre, N_post, R) impulse responses (w_ir2 @ ibasis.T, impulse.py:65, transposed :281).
    Returns S (nT,N) float64 spike counts and X (nT,N) activations."""
    N = bias.shape[0]
    T_imp = imps.shape[2]
    X = np.tile(bias[None, :], (nT, 1)).astype(np.float64)
    S = np.zeros((nT, N))
    acc = np.zeros(N)
    thr = -np.log(rng.random(N))
    AW = (A.astype(np.float64) * W)[:, :, None] * imps      # (pre, post, R)
    max_spks_per_bin = 10
    for t in range(nT):
        lam = nlin(X[t, :], nlin_type)
        acc = acc + lam * dt
        i_spk = acc > thr
        S[t, i_spk] += 1
        n_spk = int(np.sum(i_spk))
        t_imp = min(nT - t - 1, T_imp)
        while n_spk > 0:
            if np.any(S[t, :] >= max_spks_per_bin):
                break
            X[t + 1:t + t_imp + 1, :] += np.sum(AW[i_spk, :, :t_imp], 0).T
            acc -= thr * i_spk
            acc[acc < 0] = 0
            thr[i_spk] = -np.log(rng.random(n_spk))
            i_spk = acc > thr
            S[t, i_spk] += 1
            n_spk = int(np.sum(i_spk))
    return S, X


def sample_group_lasso(rng, N, B, mu, sigma, lam):
    """GroupLasso.sample, priors.py:215-224."""
    norms = rng.laplace(0, lam, size=(N, 1))
    v = mu + sigma * rng.standard_normal((N, B))
    v_norms = np.sqrt(np.sum(v ** 2, axis=1)).reshape(N, 1)
    return v * norms / v_norms
