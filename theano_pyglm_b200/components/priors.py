"""Priors on the hot path's parameters (interface of pyglm/components/priors.py).

Gaussian (:125-158) and GroupLasso (:188-224) are O(N*B) host-side terms added to the engine's
ll / gradient; the latent-variable priors of the reference are out of scope."""
import numpy as np

from .component import Component, Shared


def create_prior(model, **kwargs):
    typ = model['type'].lower()
    if typ in ('normal', 'gaussian'):
        return Gaussian(model, **kwargs)
    if typ in ('group_lasso', 'grouplasso'):
        return GroupLasso(model, **kwargs)
    raise Exception("Unrecognized prior type: %s" % typ)


class Gaussian(Component):
    def __init__(self, model, name='gaussian'):
        self.prms = model
        self.mu = Shared(model['mu'], 'mu')
        self.sigma = Shared(model['sigma'], 'sigma')

    def log_p(self, value):
        """-0.5/sigma^2 * sum((value-mu)^2)   (priors.py:139; no normalising constant)."""
        return -0.5 / self.sigma.get_value() ** 2 * np.sum((np.asarray(value) - self.mu.get_value()) ** 2)

    def grad_log_p(self, value):
        return -(np.asarray(value) - self.mu.get_value()) / self.sigma.get_value() ** 2

    def log_p_batch(self, value):
        """log_p of every leading-axis slice of value (M, groups, B) -> (M,)."""
        v = np.asarray(value)
        return -0.5 / self.sigma.get_value() ** 2 * np.sum((v - self.mu.get_value()) ** 2, axis=tuple(range(1, v.ndim)))

    grad_log_p_batch = grad_log_p                       # elementwise

    def set_hyperparameters(self, model):
        self.mu.set_value(model['mu'])
        self.sigma.set_value(model['sigma'])

    def sample(self, acc, size=(1,)):
        return self.mu.get_value() + self.sigma.get_value() * np.random.randn(*size)


class GroupLasso(Component):
    def __init__(self, model, name='gaussian'):
        self.prms = model
        self.lam = Shared(model['lam'], 'lam')
        self.mu = Shared(model['mu'], 'mu')
        self.sigma = Shared(model['sigma'], 'sigma')

    def _z(self, value):
        return (np.asarray(value) - self.mu.get_value()) / self.sigma.get_value()

    def log_p(self, value):
        """-lam * sum_groups ||(value-mu)/sigma||_2, groups = rows (priors.py:202)."""
        return -1.0 * self.lam.get_value() * np.sum(np.sqrt(np.sum(self._z(value) ** 2, axis=1)))

    def grad_log_p(self, value):
        """NaN for an all-zero group, like the symbolic gradient of sqrt at 0 in the reference;
        fit_glm's NaN guard (coord_descent.py:176-180) relies on seeing it."""
        z = self._z(value)
        nrm = np.sqrt(np.sum(z ** 2, axis=1, keepdims=True))
        with np.errstate(invalid='ignore', divide='ignore'):
            return -self.lam.get_value() * z / nrm / self.sigma.get_value()

    def log_p_batch(self, value):
        """value (M, groups, B) -> (M,): the group norms run over the last axis."""
        return -1.0 * self.lam.get_value() * np.sum(np.sqrt(np.sum(self._z(value) ** 2, axis=-1)), axis=-1)

    def grad_log_p_batch(self, value):
        z = self._z(value)
        nrm = np.sqrt(np.sum(z ** 2, axis=-1, keepdims=True))
        with np.errstate(invalid='ignore', divide='ignore'):
            return -self.lam.get_value() * z / nrm / self.sigma.get_value()

    def set_hyperparameters(self, model):
        self.mu.set_value(model['mu'])
        self.sigma.set_value(model['sigma'])
        self.lam.set_value(model['lam'])

    def sample(self, acc, size=(1,)):
        """Laplace-distributed group norms on Gaussian directions (priors.py:215-224)."""
        N = size[0]
        norms = np.random.laplace(0, self.lam.get_value(), size=(N, 1))
        v = self.mu.get_value() + self.sigma.get_value() * np.random.randn(*size)
        return v * norms / np.sqrt(np.sum(v ** 2, axis=1)).reshape(N, 1)
