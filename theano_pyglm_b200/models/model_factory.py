"""make_model / stabilize_sparsity / check_stability (interface of pyglm/models/model_factory.py)."""
import copy

import numpy as np

from .sparse_weighted_model import SparseWeightedModel
from .standard_glm import StandardGlm

_TEMPLATES = {
    'standard_glm': StandardGlm, 'standardglm': StandardGlm,
    'sparse_weighted_model': SparseWeightedModel, 'sparseweightedmodel': SparseWeightedModel,
}


def make_model(template, N=None, dt=None):
    """Deep-copy a template (by name or dict) and override N / dt (model_factory.py:18-67)."""
    if isinstance(template, str):
        key = template.lower()
        if key not in _TEMPLATES:
            raise Exception("Unrecognized template model: %s!" % template)
        model = copy.deepcopy(_TEMPLATES[key])
    elif isinstance(template, dict):
        model = copy.deepcopy(template)
    else:
        raise Exception("Unrecognized template model!")
    if N is not None:
        model['N'] = N
    if dt is not None:
        model['dt'] = dt
    return model


def stabilize_sparsity(model):
    """rho <- min(1, maxeig^2 / (N sigma^2)) for Gaussian weights on an Erdos-Renyi graph, with
    maxeig = 0.7 shifted by the refractory mean (model_factory.py:69-102)."""
    graph_model = model['network']['graph']
    weight_model = model['network']['weight']
    if graph_model['type'].lower() != 'erdos_renyi':
        return
    if weight_model.get('prior', {}).get('type', '').lower() != 'gaussian':
        return
    maxeig = 1.0 - 0.3
    if 'refractory_prior' in weight_model:
        maxeig -= weight_model['refractory_prior']['mu']
    sigma = weight_model['prior']['sigma']
    graph_model['rho'] = float(np.minimum(maxeig ** 2 / model['N'] / sigma ** 2, 1.0))


def check_stability(model, x, N):
    """Spectral abscissa of A*W below 1 (model_factory.py:173-185)."""
    if model['network']['weight']['type'].lower() == 'gaussian':
        Weff = x['net']['graph']['A'] * np.reshape(x['net']['weights']['W'], (N, N))
        return bool(np.amax(np.real(np.linalg.eigvals(Weff))) < 1)
    return True
