"""Population of coupled GLMs (interface of pyglm/population.py).

Same public surface as the reference class -- Population(model); .add_data/.set_data/
.preprocess_data; .sample(); .extract_vars(x, n); .get_variables(); .compute_ll/.compute_log_prior/
.compute_log_p(x); .eval_state(x); .set_hyperparameters(model) -- and the same state-dict layout
(population.py:149-162), but the N sequential `seval(glm.ll)` calls of population.py:80-84
become one engine call over resident data.  Extra methods `ll_grad` / `glm_ll_grad` expose the
batched likelihood + gradient the MAP and MCMC drivers consume.
"""
import numpy as np

from . import engine
from .components.latent import LatentVariables
from .components.network import Network
from .glm import Glm


class Population:
    def __init__(self, model, device=0, x_dtype=None, path="auto", time_group=None, time_sharded=False):
        """`time_sharded=True` (one process per GPU under torch.distributed; `time_group` = the process group, default
        the world): every rank adds ITS time shard of each recording (`utils.parallel_util.shard_data_by_time`), and every
        log-likelihood / gradient this object returns is the sum over the ranks' shards -- ll and its gradient are sums
        over time bins, the algebra the reference uses to add data sequences (population.py:41-43).  All ranks then see
        identical numbers, so the serial drivers (coord_descent, HMC) run unchanged and in lock-step on every rank."""
        self.model = model
        self.time_sharded = bool(time_sharded)
        self.time_group = time_group
        self.N = model['N']
        self.device = device
        self.path = path
        self.data_sequences = []
        self.latent = LatentVariables(model)
        self.network = Network(model, self.latent)
        self.glm = Glm(model, self.network, self.latent)
        if hasattr(self.glm.bkgd_model, 'device'):
            self.glm.bkgd_model.device = device                     # the stimulus projection runs on this rank's GPU
        # MCMC compares differences of log-likelihood sums against uniforms: keep X in FP64 there.
        stochastic = model['network']['graph']['type'].lower() != 'complete'
        self.x_dtype = x_dtype or ("f64" if stochastic else "f32")
        self._current = None

    # -- variables ------------------------------------------------------------------------------
    def get_variables(self):
        return {'latent': self.latent.get_variables(), 'net': self.network.get_variables(),
                'glm': self.glm.get_variables()}

    def set_hyperparameters(self, model):
        self.latent.set_hyperparameters(model)
        self.network.set_hyperparameters(model)
        self.glm.set_hyperparameters(model)

    def sample(self):
        """Draw parameters from the prior, same draw order as population.py:149-162."""
        v = {}
        v['latent'] = self.latent.sample(v)
        v['net'] = self.network.sample(v)
        v['glms'] = []
        for n in range(self.N):
            xn = self.glm.sample(v)
            xn['n'] = n
            v['glms'].append(xn)
        return v

    def extract_vars(self, vals, n):
        """Variables of the n-th GLM under key 'glm', everything else shared (population.py:164-175)."""
        out = {}
        for k, v in vals.items():
            if k == 'glms':
                out['glm'] = v[n]
            else:
                out[k] = v
        return out

    # -- data -----------------------------------------------------------------------------------
    def preprocess_data(self, data):
        """Upload the spikes and run the spike-history filter (K1) once; the handle stays resident.
        The reference stores data['fS'] here (impulse.py:130); use `get_fS(data)` to pull it back."""
        assert isinstance(data, dict), 'Data must be a dictionary'
        self.latent.preprocess_data(data)
        self.network.preprocess_data(data)
        self.glm.preprocess_data(data)
        data['_b200'] = engine.Dataset(data['S'], self.model['dt'], self.glm.imp_model.ibasis, halo=int(data.get('halo', 0)),
                                       x_dtype=self.x_dtype, device=self.device, fstim=data.get('fstim'))
        data['preprocessed'] = True
        return data

    def add_data(self, data, set_as_current_data=True):
        assert isinstance(data, dict), 'Data must be a dictionary'
        assert 'S' in data, 'Data must contain an array of spike times'
        assert isinstance(data['S'], np.ndarray), 'Spike times must be a numpy array'
        if 'preprocessed' not in data or not data['preprocessed'] or '_b200' not in data:
            data = self.preprocess_data(data)
        self.data_sequences.append(data)
        if set_as_current_data:
            self.set_data(data)

    def set_data(self, data):
        """Condition on a data sequence.  O(1): selects the resident handle (the reference re-copies
        S and fS into shared variables on every call, glm.py:99-110)."""
        assert 'preprocessed' in data and data['preprocessed'] is True, \
            'Data must be preprocessed before it can be set'
        self._current = data

    def get_fS(self, data=None):
        return self._handle(data).fS()

    def _handle(self, data=None):
        data = self._current if data is None else data
        if data is None:
            raise RuntimeError("no data sequence set")
        return data['_b200']

    # -- synthetic data -------------------------------------------------------------------------
    def simulate(self, vars, T_range, dt, stim=None, dt_stim=None, verbose=False):
        """Draw spikes from the network of coupled GLMs by time rescaling (population.py:233-389): each
        neuron integrates lam*dt and fires when the integral passes an Exp(1) threshold; a spike adds the
        weighted impulse responses A*W*imp to the future activation of its targets; at most 10 spikes per
        bin.  Returns (S, X): spike counts (nT, N) and the activation (nT, N), so that
        f_nlin(X) is the firing rate the likelihood kernels must reproduce (generate_synth_data.py:124-129).
        Host code: the recursion is sequential in time and is not part of the accelerated path."""
        T_start, T_stop = T_range
        N = self.N
        nT = len(np.arange(int(T_start / dt), int(T_stop / dt)))
        f = self.glm.nlin_model.f_nlin
        X = np.zeros((nT, N))
        for n in range(N):
            X[:, n] = self.glm.bias_model.I_bias(vars['glms'][n]['bias'])
        if self.glm.bkgd_model.n_vars:                                # stimulus current (population.py:259-270)
            tmp = {'S': np.zeros((nT, N)), 'stim': stim, 'dt_stim': dt_stim, 'T': float(T_stop - T_start)}
            self.glm.bkgd_model.preprocess_data(tmp)
            for n in range(N):
                X[:, n] += tmp['fstim'] @ self.glm.bkgd_model.weights(vars['glms'][n]['bkgd'])
        if verbose:
            print("Max background rate: %s" % str(f(np.amax(X))))
        # imps[pre, post, lag] scaled by the network: what one spike of `pre` adds to `post`
        imps = np.stack([self.glm.imp_model.impulse(vars['glms'][n]['imp']) for n in range(N)], axis=1)
        A, W = self.network.A(vars['net']), self.network.W(vars['net'])
        gain = (np.ones((N, N)) if A is None else A.astype(np.float64)) * (np.ones((N, N)) if W is None else W)
        kick = imps * gain[:, :, None]
        T_imp = kick.shape[2]
        S = np.zeros((nT, N))
        acc = np.zeros(N)
        thr = -np.log(np.random.rand(N))
        n_exceptions = 0
        max_spks_per_bin = 10
        for t in range(nT):
            acc = acc + f(X[t, :]) * dt
            i_spk = acc > thr
            S[t, i_spk] += 1
            n_spk = int(np.sum(i_spk))
            t_imp = min(nT - t - 1, T_imp)
            while n_spk > 0:
                if np.any(S[t, :] >= max_spks_per_bin):
                    n_exceptions += 1
                    break
                X[t + 1:t + t_imp + 1, :] += np.sum(kick[i_spk, :, :t_imp], axis=0).T
                acc -= thr * i_spk
                acc[acc < 0] = 0
                thr[i_spk] = -np.log(np.random.rand(n_spk))
                i_spk = acc > thr
                S[t, i_spk] += 1
                n_spk = int(np.sum(i_spk))
        if verbose:
            lam = f(X)
            print("Sampled %s spikes." % str(np.sum(S, 0)))
            print("Expected %s spikes." % str(np.trapz(lam, dt * np.arange(nT), axis=0)))
            print("Number of exceptions arising from multiple spikes per bin: %d" % n_exceptions)
        return S, X

    # -- probabilities ----------------------------------------------------------------------------
    def ll_grad(self, x, n_lo=0, n_hi=None, data=None, grad=True):
        """Per-neuron log-likelihoods (and gradients wrt the engine's dense blocks) on one sequence:
        ll (n,), g_bias (n,), g_w (n, N*B) for neurons [n_lo, n_hi)."""
        bias, w, A, W = self.glm.engine_params(x)
        return self._sum_over_time_shards(
            self._handle(data).ll_grad(bias, w, A, W, nlin=self.glm.nlin_model.code, n_lo=n_lo, n_hi=n_hi,
                                       path=self.path, grad=grad, w_stim=self.glm.stim_weights(x)))

    def _sum_over_time_shards(self, out):
        """Time-sharded populations: sum per-neuron ll / gradient blocks over the ranks (one collective per call)."""
        if not self.time_sharded:
            return out
        from .utils.parallel_util import allreduce_sum, collective_device
        single = not isinstance(out, tuple)
        parts = allreduce_sum([out] if single else list(out), device=collective_device(self.device, self.time_group),
                              group=self.time_group)
        return parts[0] if single else tuple(parts)

    def compute_ll(self, vars):
        """sum_n ll_n on the current data sequence (population.py:71-86)."""
        return float(np.sum(self.ll_grad(vars, grad=False)))

    def _ll_grad_blocks(self, data, bias, w, A, W, ws, **kw):
        """ll, g_bias, g_w, g_w_stim (zero-width when the model has no stimulus) on one sequence."""
        out = self._sum_over_time_shards(
            data['_b200'].ll_grad(bias, w, A, W, nlin=self.glm.nlin_model.code, path=self.path, w_stim=ws, **kw))
        return out if len(out) == 4 else out + (np.zeros((len(out[0]), 0)),)

    def compute_log_prior(self, vars):
        lp = 0.0
        lp += self.latent.log_p(vars.get('latent', {}))
        lp += self.network.log_p(vars['net'])
        for n in range(self.N):
            lp += self.glm.log_prior(vars['glms'][n])
        return float(lp)

    def compute_log_p(self, vars):
        """log prior + sum over data sequences of lkhd_scale * ll (population.py:34-45, glm.py:62-63)."""
        lp = self.compute_log_prior(vars)
        scale = self.glm.lkhd_scale.get_value()
        for data in self.data_sequences:
            self.set_data(data)
            lp += scale * self.compute_ll(vars)
        return lp

    def glm_log_p_grad(self, x, n):
        """log posterior of neuron n's GLM variables and its gradient as a flat vector in the
        reference's sorted-key order (bias, bkgd, imp): what coord_descent.nlp/grad_nlp evaluate
        (coord_descent.py:40-80) -- prior plus the likelihood summed over data sequences."""
        xn = x['glms'][n]
        lp = self.glm.log_prior(xn)
        gp = self.glm.grad_log_prior(xn)
        bias, w, A, W = self.glm.engine_params(x)
        ws = self.glm.stim_weights(x)
        g_bias, g_w, g_s = 0.0, 0.0, 0.0
        scale = self.glm.lkhd_scale.get_value()
        for data in self.data_sequences:
            ll, gb, gw, gs = self._ll_grad_blocks(data, bias, w, A, W, ws, n_lo=n, n_hi=n + 1)
            lp += scale * ll[0]
            g_bias = g_bias + scale * gb[0]
            g_w = g_w + scale * gw[0]
            g_s = g_s + scale * gs[0]
        g_imp = self.glm.imp_model.chain_rule(xn['imp'], g_w)
        parts = [gp['bias']['bias'] + g_bias]
        if ws is not None:
            parts.append(gp['bkgd']['w_stim'] + g_s)
        for k in sorted(g_imp):
            parts.append(gp['imp'][k] + g_imp[k])
        return float(lp), np.concatenate([np.ravel(p) for p in parts])

    def glm_param_vector(self, xn):
        """Differentiable GLM variables of one neuron as a flat vector, sorted-key order (bias < bkgd < imp)."""
        parts = [np.ravel(xn['bias']['bias'])]
        if self.glm.bkgd_model.n_vars:
            parts.append(np.ravel(xn['bkgd']['w_stim']))
        for k in sorted(self.glm.imp_model.get_variables()):
            parts.append(np.ravel(xn['imp'][k]))
        return np.concatenate(parts).astype(np.float64)

    def set_glm_param_vector(self, xn, vec):
        xn['bias']['bias'] = np.array(vec[:1], dtype=np.float64)
        off = 1
        F = self.glm.bkgd_model.n_vars
        if F:
            xn['bkgd']['w_stim'] = np.array(vec[off:off + F], dtype=np.float64)
            off += F
        for k, shp in sorted(self.glm.imp_model.get_variables().items()):
            sz = int(np.prod(shp))
            xn['imp'][k] = np.array(vec[off:off + sz], dtype=np.float64).reshape(shp)
            off += sz

    def glms_log_p_grad(self, x):
        """The per-neuron log posteriors of coord_descent.nlp/grad_nlp for ALL neurons from one engine call
        per data sequence: lp (N,), grad (N, D) in the order of `glm_param_vector`."""
        bias, w, A, W = self.glm.engine_params(x)
        ws = self.glm.stim_weights(x)
        scale = self.glm.lkhd_scale.get_value()
        ll, gb, gw, gs = 0.0, 0.0, 0.0, 0.0
        for data in self.data_sequences:
            l, b, g, s = self._ll_grad_blocks(data, bias, w, A, W, ws)
            ll, gb, gw, gs = ll + scale * l, gb + scale * b, gw + scale * g, gs + scale * s
        lps, grads = [], []
        for n in range(self.N):
            xn = x['glms'][n]
            gp = self.glm.grad_log_prior(xn)
            g_imp = self.glm.imp_model.chain_rule(xn['imp'], gw[n])
            parts = [gp['bias']['bias'] + gb[n]]
            if ws is not None:
                parts.append(gp['bkgd']['w_stim'] + gs[n])
            parts += [np.ravel(gp['imp'][k] + g_imp[k]) for k in sorted(g_imp)]
            lps.append(self.glm.log_prior(xn) + ll[n])
            grads.append(np.concatenate(parts))
        return np.array(lps), np.stack(grads)

    # -- the same, on dense arrays (no per-neuron dict traffic): what the lock-step HMC updates call ------------------
    def dense_glm_params(self, x):
        """P (N, D): row n = glm_param_vector(x['glms'][n])."""
        return np.stack([self.glm_param_vector(x['glms'][n]) for n in range(self.N)])

    def set_dense_glm_params(self, x, P, n_lo=0, n_hi=None):
        for n in range(n_lo, self.N if n_hi is None else n_hi):
            self.set_glm_param_vector(x['glms'][n], P[n])

    def glms_log_p_grad_dense(self, P, x, n_lo=0, n_hi=None):
        """glms_log_p_grad for the parameter matrix P (N, D) and the network of state x, with the priors, the
        Dirichlet normalisation and its chain rule evaluated for all neurons at once.  With a neuron range the engine
        evaluates only columns [n_lo, n_hi) (a neuron-sharded rank's share) and the results have n_hi - n_lo rows;
        P always carries every neuron (the weights of the other neurons are not read by those columns)."""
        glm = self.glm
        n_hi = self.N if n_hi is None else n_hi
        F = glm.bkgd_model.n_vars
        b = P[:, 0]
        ws = P[:, 1:1 + F] if F else None
        V = P[:, 1 + F:]
        w = glm.imp_model.batch_weights(V)
        A, W = self.network.A(x['net']), self.network.W(x['net'])
        scale = glm.lkhd_scale.get_value()
        ll, gb, gw, gs = 0.0, 0.0, 0.0, 0.0
        for data in self.data_sequences:
            l, b_, g, s_ = self._ll_grad_blocks(data, b, w, A, W, ws, n_lo=n_lo, n_hi=n_hi)
            ll, gb, gw, gs = ll + scale * l, gb + scale * b_, gw + scale * g, gs + scale * s_
        sl = slice(n_lo, n_hi)
        b, V = b[sl], V[sl]
        bm = glm.bias_model
        lp = ll - 0.5 / bm.sig_bias ** 2 * (b - bm.mu_bias) ** 2 + glm.imp_model.batch_log_p(V)
        grad = np.empty((n_hi - n_lo, P.shape[1]))
        grad[:, 0] = gb - (b - bm.mu_bias) / bm.sig_bias ** 2
        if F:
            sig = glm.bkgd_model.prior_sigma
            lp = lp + np.sum(-0.5 / sig ** 2 * ws[sl] ** 2, axis=1)
            grad[:, 1:1 + F] = gs - ws[sl] / sig ** 2
        grad[:, 1 + F:] = glm.imp_model.batch_chain_rule(V, gw) + glm.imp_model.batch_grad_log_p(V)
        return lp, grad

    def eval_state(self, vars):
        """Firing rates and currents for the current state (population.py:88-123), engine-side lam."""
        bias, w, A, W = self.glm.engine_params(vars)
        lam = self._handle().firing_rate(bias, w, A, W, nlin=self.glm.nlin_model.code,
                                         w_stim=self.glm.stim_weights(vars))
        state = {'net': {'graph': {'A': A if A is not None else np.ones((self.N, self.N))},
                         'weights': {'W': W if W is not None else np.ones((self.N, self.N))}},
                 'glms': []}
        for n in range(self.N):
            xn = vars['glms'][n]
            state['glms'].append({'lam': lam[:, n], 'I_bias': bias[n],
                                  'imp': {'impulse': self.glm.imp_model.impulse(xn['imp']),
                                          'basis': self.glm.imp_model.ibasis}})
        state['logprior'] = self.compute_log_prior(vars)
        state['ll'] = self.compute_ll(vars)
        state['logp'] = state['ll'] + state['logprior']
        return state
