"""Impulse-response components (interface of pyglm/components/impulse.py).

`preprocess_data` is where the reference runs the spike-history filter; here it hands the spike
matrix to the engine, which keeps spikes and the filtered train X resident on the GPU."""
import numpy as np

from .component import Component
from .priors import create_prior
from ..utils.basis import create_basis, interpolate_basis


def create_impulse_component(model, glm, latent):
    typ = model['impulse']['type'].lower()
    if typ == 'basis':
        return LinearBasisImpulses(model)
    if typ == 'dirichlet':
        return DirichletImpulses(model)
    raise NotImplementedError("impulse model '%s' is outside the accelerated hot path" % typ)


class _BasisImpulses(Component):
    style = "linear"

    def __init__(self, model):
        self.model = model
        self.imp_model = model['impulse']
        self.N = model['N']
        self.basis = create_basis(self.imp_model['basis'])
        self.B = self.basis.shape[1]
        self.initialize_basis()

    def initialize_basis(self):
        self.ibasis = interpolate_basis(self.basis, self.model['dt'], self.imp_model['dt_max'],
                                        self.imp_model['basis']['norm'], self.style)

    def preprocess_data(self, data):
        nT, Ns = data["S"].shape
        assert Ns == self.N, "ERROR: Spike train must be (TxN) dimensional where N=%d" % self.N

    def get_state(self):
        return {'basis': self.ibasis}


class LinearBasisImpulses(_BasisImpulses):
    """I_imp[t,pre] = sum_b ir[t,pre,b] w_ir[pre,b] (impulse.py:45-58), prior from the model dict."""
    style = "linear"

    def __init__(self, model):
        super().__init__(model)
        self.prior = create_prior(self.imp_model['prior'])

    def get_variables(self):
        return {'w_ir': (self.N * self.B,)}

    def weights(self, xn_imp):
        """(N_pre, B) block the engine consumes for this postsynaptic neuron."""
        return np.reshape(xn_imp['w_ir'], (self.N, self.B))

    def log_p(self, xn_imp):
        return self.prior.log_p(self.weights(xn_imp))                # impulse.py:60

    def grad_log_p(self, xn_imp):
        return {'w_ir': self.prior.grad_log_p(self.weights(xn_imp)).ravel()}

    def chain_rule(self, xn_imp, g_w):
        """Engine gradient wrt the (N,B) block -> gradient wrt this component's variables."""
        return {'w_ir': np.asarray(g_w).ravel()}

    def impulse(self, xn_imp):
        return self.weights(xn_imp) @ self.ibasis.T                  # impulse.py:65

    # -- all postsynaptic neurons at once: V (M, N*B) holds each neuron's variables in parameter-vector order --------
    def batch_weights(self, V):
        return np.asarray(V)

    def batch_log_p(self, V):
        return self.prior.log_p_batch(np.reshape(V, (-1, self.N, self.B)))

    def batch_grad_log_p(self, V):
        return np.reshape(self.prior.grad_log_p_batch(np.reshape(V, (-1, self.N, self.B))), np.shape(V))

    def batch_chain_rule(self, V, g_w):
        return np.asarray(g_w)

    def set_hyperparameters(self, model):
        self.prior.set_hyperparameters(model['prior'])

    def sample(self, acc):
        return {'w_ir': self.prior.sample(None, size=(self.N, self.B)).ravel()}


class DirichletImpulses(_BasisImpulses):
    """Normalised impulse responses: per presynaptic neuron a vector g_pre of B gammas and
    beta = |g| / sum|g| (impulse.py:286-291); log_p = sum (alpha-1) log|g| - |g| (:320-322)."""
    style = "dirichlet"

    def __init__(self, model):
        super().__init__(model)
        self.alpha = self.imp_model['alpha']

    def get_variables(self):
        return {'g_%d' % n: (self.B,) for n in range(self.N)}

    def _g(self, xn_imp):
        return np.stack([np.asarray(xn_imp['g_%d' % n], dtype=np.float64) for n in range(self.N)])

    def weights(self, xn_imp):
        gabs = np.abs(self._g(xn_imp))
        return gabs / gabs.sum(axis=1, keepdims=True)

    def log_p(self, xn_imp):
        gabs = np.abs(self._g(xn_imp))
        return float(np.sum((self.alpha - 1.0) * np.log(gabs) - gabs))

    def grad_log_p(self, xn_imp):
        g = self._g(xn_imp)
        gr = np.sign(g) * ((self.alpha - 1.0) / np.abs(g) - 1.0)
        return {'g_%d' % n: gr[n] for n in range(self.N)}

    def chain_rule(self, xn_imp, g_beta):
        g = self._g(xn_imp)
        s = np.abs(g).sum(axis=1, keepdims=True)
        beta = np.abs(g) / s
        g_beta = np.reshape(g_beta, (self.N, self.B))
        gr = np.sign(g) * (g_beta - np.sum(g_beta * beta, axis=1, keepdims=True)) / s
        return {'g_%d' % n: gr[n] for n in range(self.N)}

    def impulse(self, xn_imp):
        return self.weights(xn_imp) @ self.ibasis.T

    # -- all postsynaptic neurons at once.  V (M, N*B): the blocks g_k in parameter-vector order, i.e. sorted by NAME
    #    ('g_0', 'g_1', 'g_10', ...: theano_func_wrapper.py:53-67), so block j belongs to presynaptic neuron order[j].
    @property
    def order(self):
        if not hasattr(self, '_order'):
            self._order = np.array([int(name[2:]) for name in sorted(self.get_variables())])
            self._inverse = np.argsort(self._order)
        return self._order

    def _blocks(self, V):
        return np.reshape(np.asarray(V, dtype=np.float64), (-1, self.N, self.B))

    def batch_weights(self, V):
        """beta of every (post, pre) pair, pre-major like the engine's w rows: (M, N*B)."""
        g = np.abs(self._blocks(V))
        beta = g / g.sum(axis=2, keepdims=True)
        self.order
        return beta[:, self._inverse, :].reshape(len(beta), -1)

    def batch_log_p(self, V):
        g = np.abs(self._blocks(V))
        return np.sum((self.alpha - 1.0) * np.log(g) - g, axis=(1, 2))

    def batch_grad_log_p(self, V):
        g = self._blocks(V)
        return (np.sign(g) * ((self.alpha - 1.0) / np.abs(g) - 1.0)).reshape(np.shape(V))

    def batch_chain_rule(self, V, g_w):
        """d ll / d g (vector order) from the engine's d ll / d beta rows (pre-major)."""
        g = self._blocks(V)
        s = np.abs(g).sum(axis=2, keepdims=True)
        beta = np.abs(g) / s
        gb = np.reshape(g_w, (-1, self.N, self.B))[:, self.order, :]
        return (np.sign(g) * (gb - np.sum(gb * beta, axis=2, keepdims=True)) / s).reshape(np.shape(V))

    def sample(self, acc):
        return {'g_%d' % n: np.random.gamma(self.alpha, np.ones(self.B)) for n in range(self.N)}   # impulse.py:350
