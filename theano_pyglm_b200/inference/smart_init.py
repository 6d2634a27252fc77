"""Data-driven initialisation (interface of pyglm/inference/smart_init.py:7-99): spike-triggered-average stimulus
weights for a BasisStimulus background, then a dense graph."""
import numpy as np

from ..components.bkgd import BasisStimulus
from ..utils.sta import project_onto_basis, sta


def initialize_with_dense_graph(population, data, x0):
    if 'A' in x0['net']['graph']:
        x0['net']['graph']['A'] = np.ones_like(x0['net']['graph']['A'])


def initialize_stim_with_sta(population, data, x0, Ns=None):
    """smart_init.py:30-99, temporal (BasisStimulus) branch: w_stim[d*B:(d+1)*B] = projection of the STA of stimulus
    dimension d onto the interpolated stimulus basis.  The spatiotemporal branch is out of scope."""
    bk = population.glm.bkgd_model
    if not isinstance(bk, BasisStimulus) or data.get('stim') is None:
        return
    Ns = np.arange(population.N) if Ns is None else ([Ns] if isinstance(Ns, (int, np.integer)) else Ns)
    s = sta(data['stim'], data, bk.ibasis.shape[0], Ns=Ns)
    B = bk.ibasis.shape[1]
    for i, n in enumerate(Ns):
        sn = s[i]
        w_t = np.concatenate([np.ravel(project_onto_basis(sn[:, d], bk.ibasis)) for d in range(sn.shape[1])])
        assert w_t.size == B * sn.shape[1]
        x0['glms'][n]['bkgd']['w_stim'] = w_t


def initialize_with_data(population, data, x0, Ns=None):
    initialize_stim_with_sta(population, data, x0, Ns=Ns)
    initialize_with_dense_graph(population, data, x0)
