"""Component protocol of the population GLM (interface of pyglm/components/component.py:1-38).

Components here are plain NumPy/host objects: they own hyper-parameters, priors, sampling and the
mapping from the state dict to the dense parameter blocks the CUDA engine consumes.  The
likelihood itself is not assembled from component expressions (the reference builds one Theano
graph); it is evaluated by the engine.
"""


class Shared:
    """Stand-in for a Theano shared scalar (`.get_value()` / `.set_value()`), e.g. glm.lkhd_scale
    (glm.py:62) which AIS anneals through set_value (parallel_ais.py:108-109)."""

    def __init__(self, value, name=None):
        self.value, self.name = value, name

    def get_value(self):
        return self.value

    def set_value(self, value):
        self.value = value


class Component(object):
    def get_variables(self):
        """Names -> shapes of the free variables of this component."""
        return {}

    def get_state(self):
        return {}

    def preprocess_data(self, data):
        pass

    def set_data(self, data):
        pass

    def set_hyperparameters(self, model):
        pass

    def sample(self, acc):
        """Draw the variables from the prior."""
        return {}
