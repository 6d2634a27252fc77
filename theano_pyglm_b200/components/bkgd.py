"""Background / stimulus component (interface of pyglm/components/bkgd.py).

Every north-star configuration uses `bkgd: none` (models/standard_glm.py:27,
sparse_weighted_model.py:27); `BasisStimulus` (bkgd.py:45-172) is the stimulus row of the hot path:
its filtered stimulus rides in the engine's feature matrix behind the spike-history features, so
I_stim = fstim @ w_stim and d ll / d w_stim come out of the same fused kernels.
The spatiotemporal / shared-tuning-curve variants are out of scope."""
import numpy as np

from .. import engine
from ..utils.basis import create_basis
from .component import Component


def create_bkgd_component(model, glm, latent):
    typ = model['bkgd']['type'].lower()
    if typ in ('no_stimulus', 'none', 'nostimulus'):
        return NoStimulus(model)
    if typ == 'basis':
        return BasisStimulus(model)
    raise NotImplementedError("background model '%s' is outside the accelerated hot path "
                              "(only 'none' and 'basis' are built; see DESIGN.md scope)" % typ)


class NoStimulus(Component):
    """I_stim = 0, log_p = 0 (bkgd.py:29-43)."""
    n_vars = 0

    def __init__(self, model):
        self.model = model

    def weights(self, xn):
        return np.zeros(0)

    def log_p(self, xn):
        return 0.0

    def grad_log_p(self, xn):
        return {}


class BasisStimulus(Component):
    """Stimulus filtered by a temporal basis, one weight per (stimulus dimension, basis function)."""
    prior_sigma = 0.01                                     # hard-coded in the reference (bkgd.py:76)

    def __init__(self, model, device=0):
        self.model = model
        self.bkgd_model = model['bkgd']
        self.device = device
        self.basis = create_basis(self.bkgd_model['basis'])
        self.initialize_basis()
        self.n_vars = self.ibasis.shape[1] * self.bkgd_model['D_stim']

    def initialize_basis(self):
        """bkgd.py:102-121: linspace resampling at the bin width; `norm` divides by the column SUM."""
        L, B = self.basis.shape
        Lt_int = int(round(self.bkgd_model['dt_max'] / self.model['dt']))
        t_int, t_bas = np.linspace(0, 1, Lt_int), np.linspace(0, 1, L)
        ibasis = np.stack([np.interp(t_int, t_bas, self.basis[:, b]) for b in range(B)], axis=1)
        if self.bkgd_model['basis']['norm']:
            ibasis = ibasis / np.sum(ibasis, axis=0)[None, :]
        self.ibasis = ibasis

    def get_variables(self):
        return {'w_stim': (self.n_vars,)}

    def get_state(self, xn=None):
        st = {'basis': self.ibasis}
        if xn is not None:                                 # bkgd.py:84: stim_resp = ibasis @ w_stim
            st['stim_response'] = self.ibasis @ np.asarray(xn['w_stim']).reshape(-1, self.ibasis.shape[1]).T
        return st

    def weights(self, xn):
        return np.asarray(xn['w_stim'], dtype=np.float64).reshape(-1)

    def log_p(self, xn):
        return float(np.sum(-0.5 / self.prior_sigma ** 2 * (self.weights(xn) - 0.0) ** 2))      # bkgd.py:76

    def grad_log_p(self, xn):
        return {'w_stim': -self.weights(xn) / self.prior_sigma ** 2}

    def preprocess_data(self, data):
        """bkgd.py:122-154: interpolate the stimulus onto the spike bins, project it on the basis (on the
        GPU: engine.filter_dense) and store data['fstim'] (T, D*B), d-major."""
        if not abs(data['stim'].shape[0] * data['dt_stim'] - data['T']) < data['dt_stim']:
            raise Exception('Stimulus length is not the same as data time length!')
        D = self.bkgd_model['D_stim']
        if not D == data['stim'].shape[1]:
            raise Exception("Stim dimension (%d) is not equal to that specified by model (%d)"
                            % (data['stim'].shape[1], D))
        dt, dt_stim = self.model['dt'], self.bkgd_model['dt_stim']
        t = dt * np.arange(data['S'].shape[0])
        t_stim = dt_stim * np.arange(data['stim'].shape[0])
        stim = np.stack([np.interp(t, t_stim, data['stim'][:, d]) for d in range(D)], axis=1)
        cstim = engine.filter_dense(stim, self.ibasis, device=self.device)
        data['fstim'] = cstim.reshape(cstim.shape[0], -1)

    def sample(self, acc):
        return {'w_stim': 0.01 * np.random.randn(self.n_vars)}                                   # bkgd.py:165-170
