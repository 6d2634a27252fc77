"""Nonlinearity (interface of pyglm/components/nlin.py)."""
import numpy as np

from .component import Component
from ..engine import NLIN_EXP, NLIN_SOFTPLUS


def create_nlin_component(model):
    typ = model['nonlinearity']['type'].lower()
    if typ == 'exp':
        return ExpNonlinearity(model)
    if typ == 'explinear':
        return ExpLinearNonlinearity(model)
    raise Exception("Unrecognized nonlinearity model: %s" % typ)


class ExpNonlinearity(Component):
    code = NLIN_EXP

    def __init__(self, model):
        self.f_nlin = np.exp                                         # nlin.py:29

    def log_p(self, xn):
        return 0.0


class ExpLinearNonlinearity(Component):
    """Named 'explinear' in the model dicts but evaluates log(1+exp(x)) (nlin.py:42-47)."""
    code = NLIN_SOFTPLUS

    def __init__(self, model):
        self.f_nlin = lambda x: np.logaddexp(0.0, x)

    def log_p(self, xn):
        return 0.0
