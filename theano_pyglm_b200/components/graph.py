"""Adjacency-matrix models (interface of pyglm/components/graph.py): complete and Erdos-Renyi."""
import numpy as np

from .component import Component, Shared


def create_graph_component(model, latent):
    typ = model['network']['graph']['type'].lower()
    if typ == 'complete':
        return CompleteGraphModel(model)
    if typ in ('erdos_renyi', 'erdosrenyi'):
        return ErdosRenyiGraphModel(model)
    raise NotImplementedError("graph model '%s' is outside the accelerated hot path" % typ)


class CompleteGraphModel(Component):
    def __init__(self, model):
        self.model = model

    def A(self, x_graph):
        return None                                                 # engine: NULL == all ones (graph.py:35)

    def log_p(self, x_graph):
        return 0.0


class ErdosRenyiGraphModel(Component):
    def __init__(self, model):
        self.model = model
        self.prms = model['network']['graph']
        N = model['N']
        self.rho = self.prms['rho'] * np.ones((N, N))
        if 'rho_refractory' in self.prms:
            self.rho[np.diag_indices(N)] = self.prms['rho_refractory']    # graph.py:55-56
        self.pA = Shared(self.rho, 'pA')
        self.lkhd_scale = Shared(1.0, 'lkhd_scale')

    def get_variables(self):
        N = self.model['N']
        return {'A': (N, N)}

    def A(self, x_graph):
        return x_graph['A']

    def log_p(self, x_graph):
        A = np.asarray(x_graph['A'], dtype=np.float64)
        rho = self.rho
        lk = np.sum(A * np.log(np.minimum(1.0 - 1e-8, rho)) + (1 - A) * np.log(np.maximum(1e-8, 1.0 - rho)))
        return self.lkhd_scale.get_value() * lk                      # graph.py:68-71

    def sample(self, acc):
        N = self.model['N']
        return {'A': (np.random.rand(N, N) < self.rho).astype(np.int8)}
