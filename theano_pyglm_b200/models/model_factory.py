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


def convert_model(from_popn, from_model, from_vars, to_popn, to_model, to_vars):
    """Project a fitted standard GLM (basis impulses) onto a weighted model with normalised (Dirichlet)
    impulses, to start MCMC from the MAP estimate (model_factory.py:187-268, used at gibbs.py:2490-2507).
    Each impulse response is fitted by non-negative least squares on the target basis with either sign; its
    area becomes W[n1,n2], its shape the Dirichlet block g_{n1}; A keeps the strongest ~2*rho of the edges."""
    import copy
    from scipy.optimize import nnls
    N = from_popn.N
    if not (from_model['impulse']['type'].lower() == 'basis' and
            to_model['impulse']['type'].lower() == 'dirichlet'):
        raise NotImplementedError("convert_model: only basis -> dirichlet impulses are built")
    conv = copy.deepcopy(to_vars)
    basis = to_popn.glm.imp_model.ibasis                                  # (R, B)
    alpha, B = to_popn.glm.imp_model.alpha, to_popn.glm.imp_model.B
    W = np.zeros((N, N))
    for n2 in range(N):
        imp = from_popn.glm.imp_model.impulse(from_vars['glms'][n2]['imp'])      # (N_pre, R)
        for n1 in range(N):
            wp, rp = nnls(basis, imp[n1, :])
            wn, rn = nnls(basis, -1.0 * imp[n1, :])
            sgn, w = (1.0, wp) if rp < rn else (-1.0, wn)
            w = np.clip(w, 0.001, np.inf)
            W[n1, n2] = sgn * np.sum(w)
            conv['glms'][n2]['imp']['g_%d' % n1] = alpha * B * w / np.sum(w)
    conv['net']['weights']['W'] = W.flatten()
    if 'rho' in to_model['network']['graph']:
        W_sorted = np.sort(np.abs(W.ravel()))
        k = int(np.floor((1.0 - 2.0 * to_model['network']['graph']['rho']) * (N ** 2 - N) - N))
        thresh = W_sorted[max(k, 0)]
        conv['net']['graph']['A'] = (np.abs(W) >= thresh).astype(np.int8)
    else:
        conv['net']['graph']['A'] = np.ones((N, N), dtype=np.int8)
    for n in range(N):
        conv['glms'][n]['bias']['bias'] = from_vars['glms'][n]['bias']['bias']
    return conv
