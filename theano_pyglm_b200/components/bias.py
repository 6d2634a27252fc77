"""Scalar bias with a Gaussian prior (interface of pyglm/components/bias.py)."""
import numpy as np

from .component import Component


def create_bias_component(model, glm, latent):
    typ = model['bias']['type'].lower()
    if typ == 'constant':
        return ConstantBias(model)
    raise Exception("Unrecognized bias model: %s" % typ)


class ConstantBias(Component):
    def __init__(self, model):
        self.mu_bias = model['bias']['mu']
        self.sig_bias = model['bias']['sigma']

    def get_variables(self):
        return {'bias': (1,)}

    def I_bias(self, xn):
        return xn['bias'][0]                                         # bias.py:32

    def log_p(self, xn):
        return -0.5 / self.sig_bias ** 2 * (xn['bias'][0] - self.mu_bias) ** 2      # bias.py:33

    def grad_log_p(self, xn):
        return {'bias': np.array([-(xn['bias'][0] - self.mu_bias) / self.sig_bias ** 2])}

    def set_hyperparameters(self, model):
        self.mu_bias, self.sig_bias = model['mu'], model['sigma']

    def sample(self, acc):
        return {'bias': self.mu_bias + self.sig_bias * np.random.randn(1,)}          # bias.py:51-56
