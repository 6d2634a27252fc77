"""Background / stimulus component (interface of pyglm/components/bkgd.py).

Every north-star configuration uses `bkgd: none` (models/standard_glm.py:27,
sparse_weighted_model.py:27).  BasisStimulus is a "next" row."""
from .component import Component


def create_bkgd_component(model, glm, latent):
    typ = model['bkgd']['type'].lower()
    if typ in ('no_stimulus', 'none', 'nostimulus'):
        return NoStimulus(model)
    raise NotImplementedError("background model '%s' is outside the accelerated hot path "
                              "(only 'none' is built; see DESIGN.md scope)" % typ)


class NoStimulus(Component):
    """I_stim = 0, log_p = 0 (bkgd.py:29-43)."""

    def __init__(self, model):
        self.model = model

    def I_stim(self, xn):
        return 0.0

    def log_p(self, xn):
        return 0.0

    def grad_log_p(self, xn):
        return {}
