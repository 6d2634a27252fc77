"""Model template: Erdos-Renyi graph, Gaussian weights, Dirichlet impulses (schema of
pyglm/models/sparse_weighted_model.py)."""
SparseWeightedModel = {
    'N': 1,
    'nonlinearity': {'type': 'explinear'},
    'bias': {'type': 'constant', 'mu': 20.0, 'sigma': 0.25},
    'bkgd': {
        'type': 'no_stimulus', 'D_stim': 1, 'dt_max': 0.3, 'mu': 0, 'sigma': 0.5,
        'basis': {'type': 'cosine', 'n_eye': 0, 'n_cos': 3, 'a': 1.0 / 120, 'b': 0.5, 'orth': False, 'norm': True},
    },
    'impulse': {
        'type': 'dirichlet', 'dt_max': 0.2, 'alpha': 1,
        'basis': {'type': 'cosine', 'n_eye': 0, 'n_cos': 5, 'a': 1.0 / 120, 'b': 0.5, 'orth': False, 'norm': True},
    },
    'network': {
        'weight': {'type': 'gaussian',
                   'prior': {'type': 'gaussian', 'mu': 0.0, 'sigma': 1.0},
                   'refractory_prior': {'type': 'gaussian', 'mu': -0.2, 'sigma': 0.5}},
        'graph': {'type': 'erdos_renyi', 'rho': 0.5, 'rho_refractory': 1.0},
    },
}
