"""Weight-matrix models (interface of pyglm/components/weights.py): constant and Gaussian."""
import numpy as np

from .component import Component
from .priors import create_prior


def create_weight_component(model, latent):
    typ = model['network']['weight']['type'].lower()
    if typ == 'constant':
        return ConstantWeightModel(model)
    if typ == 'gaussian':
        return GaussianWeightModel(model)
    raise Exception("Unrecognized weight model: %s" % typ)


class ConstantWeightModel(Component):
    def __init__(self, model):
        self.model = model
        self.value = model['network']['weight']['value']

    def W(self, x_weights):
        N = self.model['N']
        return None if self.value == 1.0 else self.value * np.ones((N, N))     # weights.py:32

    def log_p(self, x_weights):
        return 0.0


class GaussianWeightModel(Component):
    """W is a flat (N*N,) vector in the state dict (weights.py:64-65); the diagonal has its own
    'refractory' prior when the model provides one (weights.py:47-71)."""

    def __init__(self, model):
        self.model = model
        prms = model['network']['weight']
        self.prior = create_prior(prms['prior'])
        if 'refractory_prior' in prms:
            self.refractory_prior = create_prior(prms['refractory_prior'])

    def get_variables(self):
        N = self.model['N']
        return {'W': (N * N,)}

    def W(self, x_weights):
        N = self.model['N']
        return np.reshape(x_weights['W'], (N, N))

    def log_p(self, x_weights):
        W = self.W(x_weights)
        if hasattr(self, 'refractory_prior'):
            diag = np.eye(W.shape[0], dtype=bool)
            return self.prior.log_p(W[~diag]) + self.refractory_prior.log_p(W[diag])
        return self.prior.log_p(W)

    def sample(self, acc):
        N = self.model['N']
        if hasattr(self, 'refractory_prior'):                        # weights.py:78-86: diagonal first
            W = np.zeros((N, N))
            W_diag = self.refractory_prior.sample(None, (N,))
            W_off = self.prior.sample(None, (N ** 2 - N,))
            diag = np.eye(N, dtype=bool)
            W[diag] = W_diag
            lower = np.tril_indices(N, k=-1)
            upper = np.triu_indices(N, k=1)
            nl = len(lower[0])
            W[lower] = W_off[:nl]
            W[upper] = W_off[nl:]
            return {'W': W.reshape(N ** 2)}
        return {'W': self.prior.sample(None, (N ** 2,))}
