"""Network = graph model + weight model (interface of pyglm/components/network.py:4-53)."""
from .component import Component
from .graph import create_graph_component
from .weights import create_weight_component


class Network(Component):
    def __init__(self, model, latent):
        self.model = model
        self.latent = latent
        self.prms = model['network']
        self.graph = create_graph_component(model, latent)
        self.weights = create_weight_component(model, latent)

    def get_variables(self):
        return {'graph': self.graph.get_variables(), 'weights': self.weights.get_variables()}

    def log_p(self, x_net):
        return self.graph.log_p(x_net.get('graph', {})) + self.weights.log_p(x_net.get('weights', {}))

    def A(self, x_net):
        return self.graph.A(x_net.get('graph', {}))

    def W(self, x_net):
        return self.weights.W(x_net.get('weights', {}))

    def sample(self, acc):
        return {'graph': self.graph.sample(acc), 'weights': self.weights.sample(acc)}
