"""MCMC over the network GLM (interface of pyglm/inference/gibbs.py for the accelerated path).

`gibbs_sample(population, N_samples, x0, init_from_mle, callback)` keeps the reference's loop
(:2475-2571): every sweep applies, for each neuron, HMC on the bias, HMC on the impulse
responses and the collapsed spike-and-slab Gibbs update of one column of A/W.

The column update is `CollapsedGibbsNetworkColumnUpdate` (:775-1250).  Its likelihood work --
`_precompute_other_current` + 11 `_glm_ll` passes per edge + ARS probes -- is the engine's K4:
`gibbs_begin` once per sweep makes I_net resident, every edge is one `gibbs_delta_ll` call and
one rank-1 `gibbs_commit`.  Two schedules:
  * `update(x, n)`            one column at a time, edges in np.random.shuffle order, consuming
                              np.random exactly where the reference does (:1236-1237, log_sum_exp.py:26,
                              :1063) -- drop-in for seeded runs up to the ARS draws (hips is not vendored);
  * `sweep_batched(x)`        all columns in lock-step, one batched kernel launch per step of N edges --
                              columns are conditionally independent (gibbs.py:53-61), so this is the
                              same Markov kernel the reference runs in parallel over IPython engines.
"""
import copy

import numpy as np
from scipy.special import logsumexp

from ..components.impulse import DirichletImpulses
from .ars import adaptive_rejection_sample
from .hmc import hmc, hmc_batched
from .log_sum_exp import log_sum_exp_sample


class MetropolisHastingsUpdate(object):
    def preprocess(self, population):
        self.population = population

    def update(self, x):
        return x


class ParallelMetropolisHastingsUpdate(MetropolisHastingsUpdate):
    """Updates that act on one neuron given the rest; conditionally independent across n (gibbs.py:53-61)."""

    def update(self, x, n):
        return x


# ------------------------------------------------------------------------------------------------
# HMC updates of the GLM parameters (gibbs.py:164-321, :449-772); likelihood + gradient from the engine
# ------------------------------------------------------------------------------------------------
class _HmcGlmBlockUpdate(ParallelMetropolisHastingsUpdate):
    n_steps = 10

    def __init__(self):
        self.avg_accept_rate = 0.9
        self.step_sz = 0.1

    def _logp_and_grad(self, x, n):
        return self.population.glm_log_p_grad(x, n)

    def _run(self, x, n, get, put, sl):
        """HMC on the slice `sl` of neuron n's parameter vector."""
        popn = self.population
        xn = x['glms'][n]
        full0 = popn.glm_param_vector(xn)

        def with_block(q):
            v = full0.copy()
            v[sl] = q
            popn.set_glm_param_vector(xn, v)

        def U(q):
            with_block(q)
            return -self._logp_and_grad(x, n)[0]

        def grad_U(q):
            with_block(q)
            return -self._logp_and_grad(x, n)[1][sl]

        q, self.step_sz, self.avg_accept_rate = hmc(U, grad_U, self.step_sz, self.n_steps, full0[sl],
                                                    adaptive_step_sz=True, avg_accept_rate=self.avg_accept_rate)
        with_block(q)
        return x


class HmcBiasUpdate(_HmcGlmBlockUpdate):
    """gibbs.py:164-321: 10 leapfrog steps on the scalar bias."""
    n_steps = 10

    def update(self, x, n):
        return self._run(x, n, None, None, slice(0, 1))


class HmcBkgdUpdate(_HmcGlmBlockUpdate):
    """gibbs.py:324-446: HMC on w_stim; nothing to sample for the `none` background model.  (The reference's
    gradient drops the prior term, :400 -- see SURVEY appendix A; here the full log posterior is used.)"""
    n_steps = 10

    def update(self, x, n):
        F = self.population.glm.bkgd_model.n_vars
        return self._run(x, n, None, None, slice(1, 1 + F)) if F else x


class HmcImpulseUpdate(_HmcGlmBlockUpdate):
    """gibbs.py:449-566: HMC on w_ir of a LinearBasisImpulses model."""
    n_steps = 10

    def update(self, x, n):
        o = 1 + self.population.glm.bkgd_model.n_vars         # vector order: bias, w_stim, w_ir
        return self._run(x, n, None, None, slice(o, o + self.population.N * self.population.glm.imp_model.B))


class HmcDirichletImpulseUpdate(_HmcGlmBlockUpdate):
    """gibbs.py:569-772: per presynaptic neuron, 2 leapfrog steps on g_{n_pre} where an edge exists,
    a Gamma(alpha, 1) prior draw where it does not (:764-767)."""
    n_steps = 2

    def update(self, x, n_post):
        popn = self.population
        imp = popn.glm.imp_model
        A = x['net']['graph']['A']
        names = sorted(imp.get_variables())                   # vector order of the g blocks
        for n_pre in range(popn.N):
            key = 'g_%d' % n_pre
            if A[n_pre, n_post]:
                k = names.index(key)
                o = 1 + popn.glm.bkgd_model.n_vars
                self._run(x, n_post, None, None, slice(o + k * imp.B, o + (k + 1) * imp.B))
            else:
                x['glms'][n_post]['imp'][key] = np.random.gamma(imp.alpha, np.ones(imp.B))
        return x


# ------------------------------------------------------------------------------------------------
# The same HMC updates for all neurons in lock-step (one engine call per leapfrog step)
# ------------------------------------------------------------------------------------------------
class BatchedHmcGlmUpdate(MetropolisHastingsUpdate):
    """HMC on one block of every neuron's GLM parameters at once.  The chains are the reference's per-neuron chains
    (gibbs.py:164-321 bias, :324-446 background, :449-566 impulse weights, :569-772 Dirichlet impulses): same
    leapfrog, same accept rule, per-neuron adaptive step sizes; they only share their engine calls, which is what
    the conditional independence of the GLMs (gibbs.py:53-61) allows."""

    def __init__(self, block, n_steps):
        self.block, self.n_steps = block, n_steps
        self.step_sz = None
        self.avg_accept_rate = None

    def _slices(self, x):
        """(lo, hi, active) for this block: columns [lo, hi) of the parameter vectors, active (N, hi-lo)."""
        popn = self.population
        N, F = popn.N, popn.glm.bkgd_model.n_vars
        imp = popn.glm.imp_model
        if self.block == 'bias':
            return 0, 1, np.ones((N, 1))
        if self.block == 'bkgd':
            return 1, 1 + F, np.ones((N, F))
        lo = 1 + F
        if self.block == 'imp':
            return lo, lo + N * imp.B, np.ones((N, N * imp.B))
        k = self.block[1]                                  # ('g', k): Dirichlet block of presynaptic neuron k
        names = sorted(imp.get_variables())
        j = names.index('g_%d' % k)
        A = x['net']['graph']['A']
        return lo + j * imp.B, lo + (j + 1) * imp.B, np.repeat(A[k, :].astype(np.float64)[:, None], imp.B, axis=1)

    def update_dense(self, P, x, n_lo=0, n_hi=None):
        """One HMC transition of this block for neurons [n_lo, n_hi), in place on the parameter matrix P (N, D)."""
        popn = self.population
        N = popn.N
        n_hi = N if n_hi is None else n_hi
        lo, hi, active = self._slices(x)
        if hi == lo:
            return P
        own = np.zeros((N, 1))
        own[n_lo:n_hi] = 1.0                               # a neuron-sharded rank moves only its own neurons
        active = active * own
        if not active.any():
            return P
        if self.step_sz is None:
            self.step_sz = np.full(N, 0.1)
            self.avg_accept_rate = np.full(N, 0.9)

        def U_and_grad(Q):
            Pq = P.copy()
            Pq[n_lo:n_hi, lo:hi] = Q[n_lo:n_hi]
            lp, g = popn.glms_log_p_grad_dense(Pq, x, n_lo, n_hi)         # the engine evaluates only this shard's columns
            U, gU = np.zeros(N), np.zeros((N, hi - lo))
            U[n_lo:n_hi], gU[n_lo:n_hi] = -lp, -g[:, lo:hi]
            return U, gU

        q, self.step_sz, self.avg_accept_rate = hmc_batched(U_and_grad, self.step_sz, self.n_steps, P[:, lo:hi],
                                                            active=active, avg_accept_rate=self.avg_accept_rate)
        P[n_lo:n_hi, lo:hi] = q[n_lo:n_hi]
        return P

    def update(self, x, n_lo=0, n_hi=None):
        P = self.update_dense(self.population.dense_glm_params(x), x, n_lo, n_hi)
        self.population.set_dense_glm_params(x, P, n_lo, n_hi)
        return x


class BatchedDirichletImpulseUpdate(MetropolisHastingsUpdate):
    """gibbs.py:569-772 for all postsynaptic neurons at once: for each presynaptic k, two leapfrog steps on g_k in
    the neurons that have the edge k -> n, a Gamma(alpha, 1) prior draw in those that do not (:764-767)."""

    def preprocess(self, population):
        self.population = population
        self.blocks = [BatchedHmcGlmUpdate(('g', k), 2) for k in range(population.N)]
        for b in self.blocks:
            b.preprocess(population)

    def update(self, x, n_lo=0, n_hi=None):
        popn = self.population
        imp = popn.glm.imp_model
        n_hi = popn.N if n_hi is None else n_hi
        A = x['net']['graph']['A']
        P = popn.dense_glm_params(x)
        for k, blk in enumerate(self.blocks):
            blk.update_dense(P, x, n_lo, n_hi)
            lo, hi, _ = blk._slices(x)
            for n in range(n_lo, n_hi):
                if not A[k, n]:
                    P[n, lo:hi] = np.random.gamma(imp.alpha, np.ones(imp.B))
        popn.set_dense_glm_params(x, P, n_lo, n_hi)
        return x


def initialize_batched_updates(population):
    """The GLM-parameter updates of `initialize_updates` in their lock-step form."""
    ups = [BatchedHmcGlmUpdate('bias', 10), BatchedHmcGlmUpdate('bkgd', 10)]
    ups.append(BatchedDirichletImpulseUpdate() if isinstance(population.glm.imp_model, DirichletImpulses)
               else BatchedHmcGlmUpdate('imp', 10))
    for u in ups:
        u.preprocess(population)
    return ups


# ------------------------------------------------------------------------------------------------
# Collapsed Gibbs over one column of A / W
# ------------------------------------------------------------------------------------------------
class _ResidentSequences:
    """The Gibbs state of every data sequence of the population, driven as one: candidate log-likelihoods are summed
    over the sequences (gibbs.py:931-935 loops `for data in self.population.data_sequences`), commits go to all."""

    def __init__(self, handles):
        self.handles = list(handles)

    def gibbs_delta_ll(self, cols, pres, w_cand):
        out = self.handles[0].gibbs_delta_ll(cols, pres, w_cand)
        for h in self.handles[1:]:
            out = out + h.gibbs_delta_ll(cols, pres, w_cand)
        return out

    def gibbs_commit(self, cols, pres, a_new, w_new):
        for h in self.handles:
            h.gibbs_commit(cols, pres, a_new, w_new)

    def gibbs_end(self):
        for h in self.handles:
            h.gibbs_end()


class CollapsedGibbsNetworkColumnUpdate(ParallelMetropolisHastingsUpdate):
    def __init__(self):
        self.DEG_GAUSS_HERMITE = 10
        self.GAUSS_HERMITE_ABSCISSAE, self.GAUSS_HERMITE_WEIGHTS = \
            np.polynomial.hermite.hermgauss(self.DEG_GAUSS_HERMITE)        # gibbs.py:787-789
        self.sample_w_with_ars = True
        self._resident = None

    def preprocess(self, population):
        self.population = population
        self.network = population.network
        w = self.network.weights
        self.mu_w, self.sigma_w = w.prior.mu.get_value(), w.prior.sigma.get_value()
        if hasattr(w, 'refractory_prior'):
            self.mu_w_ref = w.refractory_prior.mu.get_value()
            self.sigma_w_ref = w.refractory_prior.sigma.get_value()
        else:
            self.mu_w_ref, self.sigma_w_ref = self.mu_w, self.sigma_w

    # -- engine residency ------------------------------------------------------------------------
    def begin(self, x, n_lo=0, n_hi=None):
        """Upload the state and build I_net for the columns [n_lo, n_hi) (seval(glm.I_net), gibbs.py:812-864) on EVERY
        data sequence of the population -- the conditional of A/W is the product of the sequences' likelihoods, as the
        HMC updates of the same sweep already assume; a neuron-sharded rank makes only its own columns resident."""
        popn = self.population
        bias, w, A, W = popn.glm.engine_params(x)
        seqs = popn.data_sequences if popn.data_sequences else [popn._current]
        handles = [popn._handle(d) for d in seqs]
        for h in handles:
            h.gibbs_begin(bias, w, A, W, nlin=popn.glm.nlin_model.code, n_lo=n_lo, n_hi=n_hi,
                          w_stim=popn.glm.stim_weights(x))
        self._resident = _ResidentSequences(handles)
        return self._resident

    def end(self):
        if self._resident is not None:
            self._resident.gibbs_end()
        self._resident = None

    def _prior(self, n_pre, n_post):
        if n_pre == n_post:
            return self.mu_w_ref, self.sigma_w_ref                             # :984-989
        return self.mu_w, self.sigma_w

    def _log_odds(self, log_L, ll_noA, p_A):
        """(log_pr_noA, log_pr_A) from the 10 quadrature lls and the w=0 ll (:1002-1035)."""
        log_L = np.array(log_L, dtype=np.float64)
        log_L[np.isnan(log_L)] = -np.inf
        with np.errstate(divide='ignore'):
            weighted = log_L + np.log(self.GAUSS_HERMITE_WEIGHTS / np.sqrt(np.pi))
        weighted[np.isnan(weighted)] = -np.inf
        log_G = logsumexp(weighted)
        if not np.isfinite(log_G):
            raise Exception("log_G not finie")
        with np.errstate(divide='ignore'):
            log_pr_A = np.log(p_A) + log_G
            log_pr_noA = np.log(1.0 - p_A) + ll_noA
        if np.isnan(log_pr_noA):
            log_pr_noA = -np.inf
        return log_pr_noA, log_pr_A

    def _decide_batch(self, ll, pA, u):
        """The decision rule of _collapsed_sample_AW (gibbs.py:1002-1039) + log_sum_exp_sample (log_sum_exp.py:4-37)
        for M edges at once: ll (M, 11) candidate log-likelihoods (10 quadrature nodes, then w = 0), pA (M,) prior edge
        probabilities, u (M,) uniforms.  Returns A (M,) int8: 0 where u <= p(no edge), as the reference's cumsum rule."""
        log_L = np.where(np.isnan(ll[:, :10]), -np.inf, ll[:, :10])
        with np.errstate(divide='ignore'):
            log_G = logsumexp(log_L + np.log(self.GAUSS_HERMITE_WEIGHTS / np.sqrt(np.pi))[None, :], axis=1)
            if not np.all(np.isfinite(log_G)):
                raise Exception("log_G not finie")
            lp_A = np.log(pA) + log_G
            lp_no = np.log(1.0 - pA) + ll[:, 10]
        lp_no = np.where(np.isnan(lp_no), -np.inf, lp_no)
        tot = np.logaddexp(lp_no, lp_A)
        if not np.all(np.isfinite(tot)):
            raise Exception("Total probability is zero")                         # log_sum_exp.py:20-22
        a_new = (u > np.exp(lp_no - tot)).astype(np.int8)                        # log_sum_exp.py:26-32
        if np.any(np.isclose(pA, 1.0) & (a_new == 0)):
            raise Exception("Sampled no self edge")
        return a_new

    def _sample_w(self, ds, n_pre, n_post, mu_w, sigma_w, W_nns, log_L):
        """W | A=1 by adaptive rejection sampling on the exact conditional (:1087-1126); every probe of
        the log posterior is a Q=1 delta-ll call."""
        log_post = -0.5 / sigma_w ** 2 * (W_nns - mu_w) ** 2 + log_L
        Z = np.amax(log_post)

        def _log_posterior(w):
            ll = ds.gibbs_delta_ll([n_post], [n_pre], np.array([[float(w)]]))[0, 0]
            return -0.5 / sigma_w ** 2 * (w - mu_w) ** 2 + ll - Z

        valid = np.isfinite(log_post) & (log_post > -1e8)                       # effective behaviour of :1118-1120
        return adaptive_rejection_sample(_log_posterior, W_nns[valid], log_post[valid] - Z, (-np.inf, np.inf),
                                         stepsz=sigma_w / 2.0, debug=False)

    def _resample_edge(self, ds, x, n_pre, n_post, ll, p_A):
        """Given the 11 candidate lls of one edge, draw A then W and commit (gibbs.py:1036-1066)."""
        A = x['net']['graph']['A']
        N = A.shape[0]
        W = x['net']['weights']['W'].reshape(N, N)
        mu_w, sigma_w = self._prior(n_pre, n_post)
        W_nns = np.sqrt(2) * sigma_w * self.GAUSS_HERMITE_ABSCISSAE + mu_w
        lp_noA, lp_A = self._log_odds(ll[:10], ll[10], p_A[n_pre, n_post])
        A[n_pre, n_post] = log_sum_exp_sample([lp_noA, lp_A])                   # one np.random.rand()
        if np.allclose(p_A[n_pre, n_post], 1.0) and not A[n_pre, n_post]:
            raise Exception("Sampled no self edge")
        if A[n_pre, n_post] == 1 and self.sample_w_with_ars:
            W[n_pre, n_post] = self._sample_w(ds, n_pre, n_post, mu_w, sigma_w, W_nns, np.asarray(ll[:10]))
        else:
            W[n_pre, n_post] = mu_w + sigma_w * np.random.randn()               # :1063
        x['net']['weights']['W'] = W.ravel()
        ds.gibbs_commit([n_post], [n_pre], [A[n_pre, n_post]], [W[n_pre, n_post]])

    def _candidates(self, n_pre, n_post):
        mu_w, sigma_w = self._prior(n_pre, n_post)
        return np.concatenate([np.sqrt(2) * sigma_w * self.GAUSS_HERMITE_ABSCISSAE + mu_w, [0.0]])

    # -- reference schedule: one column ----------------------------------------------------------
    def update(self, x, n):
        ds = self._resident or self.begin(x)
        N = x['net']['graph']['A'].shape[0]
        p_A = self.network.graph.pA.get_value()
        order = np.arange(N)
        np.random.shuffle(order)                                                # :1236-1237
        for n_pre in order:
            ll = ds.gibbs_delta_ll([n], [n_pre], self._candidates(n_pre, n)[None, :])[0]
            self._resample_edge(ds, x, n_pre, n, ll, p_A)
        return x

    # -- B200 schedule: all columns in lock-step ---------------------------------------------------
    def sweep_batched(self, x, n_lo=0, n_hi=None):
        """Columns [n_lo, n_hi) (all by default) advance through their shuffled presynaptic orders together."""
        N = x['net']['graph']['A'].shape[0]
        n_hi = N if n_hi is None else n_hi
        ds = self._resident or self.begin(x, n_lo, n_hi)
        p_A = self.network.graph.pA.get_value()
        cols = np.arange(n_lo, n_hi, dtype=np.int32)
        orders = np.stack([np.random.permutation(N) for _ in cols])            # one shuffled order per column
        A = x['net']['graph']['A']
        W = x['net']['weights']['W'].reshape(N, N)
        for s in range(N):
            pres = orders[:, s].astype(np.int32)
            diag = pres == cols
            mu = np.where(diag, self.mu_w_ref, self.mu_w)
            sig = np.where(diag, self.sigma_w_ref, self.sigma_w)
            W_nns = np.sqrt(2) * sig[:, None] * self.GAUSS_HERMITE_ABSCISSAE[None, :] + mu[:, None]     # (M, 10)
            ll = ds.gibbs_delta_ll(cols, pres, np.concatenate([W_nns, np.zeros((len(cols), 1))], axis=1))
            log_L = np.where(np.isnan(ll[:, :10]), -np.inf, ll[:, :10])
            pA = p_A[pres, cols]
            a_new = self._decide_batch(ll, pA, np.random.rand(len(cols)))
            w_new = mu + sig * np.random.randn(len(cols))                        # :1063 (kept where A = 0 or ARS is off)
            if self.sample_w_with_ars:
                for i in np.nonzero(a_new)[0]:                                   # W | A = 1: ARS on the exact conditional
                    w_new[i] = self._sample_w(ds, int(pres[i]), int(cols[i]), mu[i], sig[i], W_nns[i], log_L[i])
            A[pres, cols] = a_new
            W[pres, cols] = w_new
            ds.gibbs_commit(cols, pres, a_new, w_new)                           # one rank-1 update launch for the step
        x['net']['weights']['W'] = W.ravel()
        return x


# ------------------------------------------------------------------------------------------------
def initialize_updates(population):
    """gibbs.py:2413-2473 restricted to the samplers the accelerated path supports (no latent variables)."""
    serial_updates = []
    parallel_updates = [HmcBiasUpdate(), HmcBkgdUpdate()]
    parallel_updates.append(HmcDirichletImpulseUpdate() if isinstance(population.glm.imp_model, DirichletImpulses)
                            else HmcImpulseUpdate())
    parallel_updates.append(CollapsedGibbsNetworkColumnUpdate())
    for u in parallel_updates:
        u.preprocess(population)
    return serial_updates, parallel_updates


def initial_state(population, init_from_mle=False, verbose=False):
    """Prior draw, optionally moved to the MAP estimate of a standard GLM fitted to the same data and projected
    onto this model by convert_model (gibbs.py:2486-2507)."""
    x0 = population.sample()
    if init_from_mle:
        from ..models.model_factory import convert_model, make_model
        from ..population import Population
        from .coord_descent import coord_descent
        if verbose:
            print("Initializing with coordinate descent")
        mle_model = make_model('standard_glm', N=population.model['N'], dt=population.model['dt'])
        mle_popn = Population(mle_model, device=population.device)
        for data in population.data_sequences:
            mle_popn.add_data({k: v for k, v in data.items() if k not in ('_b200', 'preprocessed', 'fstim')})
        mle_x0 = coord_descent(mle_popn, x0=mle_popn.sample(), maxiter=1)
        x0 = convert_model(mle_popn, mle_model, mle_x0, population, population.model, x0)
    return x0


def gibbs_sample(population, N_samples=1000, x0=None, init_from_mle=False, callback=None, batched=True,
                 verbose=False):
    """Sample the posterior over parameters (gibbs.py:2475-2571)."""
    N = population.model['N']
    if x0 is None:
        x0 = initial_state(population, init_from_mle, verbose)
    serial_updates, parallel_updates = initialize_updates(population)
    batched_updates = initialize_batched_updates(population) if batched else None
    net_update = parallel_updates[-1]
    x = x0
    x_smpls = [copy.deepcopy(x0)]
    for smpl in range(N_samples):
        if callback is not None:
            callback(x)
        if verbose:
            print("Gibbs iteration %d. Log prob: %.3f" % (smpl, population.compute_log_p(x)))
        if batched:                                           # all neurons in lock-step: one engine call per leapfrog step
            for upd in batched_updates:
                upd.update(x)
        else:                                                 # the reference's schedule: neuron by neuron
            for upd in parallel_updates[:-1]:
                for n in range(N):
                    upd.update(x, n)
        net_update.begin(x)                                   # GLM parameters changed: rebuild the resident currents
        if batched:
            net_update.sweep_batched(x)
        else:
            for n in range(N):
                net_update.update(x, n)
        net_update.end()
        for upd in serial_updates:
            upd.update(x)
        x_smpls.append(copy.deepcopy(x))
    return x_smpls
