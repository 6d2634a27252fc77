"""One GLM shared across neurons (interface of pyglm/glm.py).

The reference builds a single symbolic graph indexed by the neuron number `n` and evaluates it
N times from Python (population.py:28-32, :80-84).  Here `Glm` only composes the components and
translates a state dict into the dense parameter blocks of the C ABI; the likelihood
glm.py:52 and its gradient are evaluated for all neurons at once by the CUDA engine.
"""
import numpy as np

from .components.bias import create_bias_component
from .components.bkgd import create_bkgd_component
from .components.component import Component, Shared
from .components.impulse import create_impulse_component
from .components.nlin import create_nlin_component


class Glm(Component):
    def __init__(self, model, network, latent):
        self.model = model
        self.network = network
        self.dt = Shared(model['dt'], 'dt')
        self.bias_model = create_bias_component(model, self, latent)
        self.bkgd_model = create_bkgd_component(model, self, latent)
        self.imp_model = create_impulse_component(model, self, latent)
        self.nlin_model = create_nlin_component(model)
        self.lkhd_scale = Shared(1.0, 'lkhd_scale')                  # glm.py:62 (AIS)

    def get_variables(self):
        return {'n': (), 'bias': self.bias_model.get_variables(), 'bkgd': self.bkgd_model.get_variables(),
                'imp': self.imp_model.get_variables(), 'nlin': self.nlin_model.get_variables()}

    def differentiable_variables(self):
        """Float-typed variables in sorted-key order (grads.py:97-117): bias, bkgd, imp, nlin."""
        v = self.get_variables()
        return {k: v[k] for k in ('bias', 'bkgd', 'imp', 'nlin')}

    def log_prior(self, xn):
        """glm.py:55-59 for one neuron's variables xn = x['glms'][n]."""
        return (self.bias_model.log_p(xn['bias']) + self.bkgd_model.log_p(xn.get('bkgd', {})) +
                self.imp_model.log_p(xn['imp']) + self.nlin_model.log_p(xn.get('nlin', {})))

    def grad_log_prior(self, xn):
        return {'bias': self.bias_model.grad_log_p(xn['bias']), 'bkgd': self.bkgd_model.grad_log_p(xn.get('bkgd', {})),
                'imp': self.imp_model.grad_log_p(xn['imp']), 'nlin': {}}

    def preprocess_data(self, data):
        self.bias_model.preprocess_data(data)
        self.bkgd_model.preprocess_data(data)
        self.imp_model.preprocess_data(data)
        self.nlin_model.preprocess_data(data)

    def set_hyperparameters(self, model):
        self.bkgd_model.set_hyperparameters(model['bkgd'])
        self.imp_model.set_hyperparameters(model['impulse'])
        self.bias_model.set_hyperparameters(model['bias'])

    def sample(self, acc):
        return {'n': -1, 'bias': self.bias_model.sample(acc), 'bkgd': self.bkgd_model.sample(acc),
                'imp': self.imp_model.sample(acc), 'nlin': self.nlin_model.sample(acc)}

    # -- state dict -> engine parameter blocks -------------------------------------------------
    def engine_params(self, x):
        """bias (N,), w (N_post, N_pre*B), A (N,N) int8 or None, W (N,N) or None from the state dict
        layout of population.py:149-162."""
        N = self.model['N']
        bias = np.array([self.bias_model.I_bias(x['glms'][n]['bias']) for n in range(N)], dtype=np.float64)
        w = np.stack([self.imp_model.weights(x['glms'][n]['imp']).reshape(-1) for n in range(N)])
        return bias, w, self.network.A(x['net']), self.network.W(x['net'])

    def stim_weights(self, x):
        """w_stim (N, F) of the state dict, or None when the model has no stimulus (bkgd.py:81)."""
        if not self.bkgd_model.n_vars:
            return None
        return np.stack([self.bkgd_model.weights(x['glms'][n]['bkgd']) for n in range(self.model['N'])])
