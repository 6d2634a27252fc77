"""Model template: fully connected GLMs with constant unit weights (schema of
pyglm/models/standard_glm.py; the nested-dict layout is the user-facing configuration API)."""
StandardGlm = {
    'N': 2,
    'nonlinearity': {'type': 'explinear'},
    'bias': {'type': 'constant', 'mu': 20, 'sigma': 0.1},
    'bkgd': {
        'type': 'none', 'D_stim': 1, 'dt_max': 0.3,
        'prior': {'type': 'spherical_gaussian', 'mu': 0.0, 'sigma': 0.01},
        'basis': {'type': 'cosine', 'n_eye': 0, 'n_cos': 3, 'a': 1.0 / 120, 'b': 0.5, 'orth': True, 'norm': False},
    },
    'impulse': {
        'type': 'basis', 'dt_max': 0.2,
        'prior': {'type': 'group_lasso', 'mu': 0.0, 'sigma': 10.0, 'lam': 1.0},
        'basis': {'type': 'cosine', 'n_eye': 0, 'n_cos': 5, 'a': 1.0 / 120, 'b': 0.5, 'orth': True, 'norm': False},
    },
    'network': {'weight': {'type': 'constant', 'value': 1.0}, 'graph': {'type': 'complete'}},
}
