"""Data-driven initialisation (interface of pyglm/inference/smart_init.py:7-18).  Only the dense-graph
initialisation applies to the models on the accelerated path (no stimulus -> no STA)."""
import numpy as np


def initialize_with_dense_graph(population, data, x0):
    if 'A' in x0['net']['graph']:
        x0['net']['graph']['A'] = np.ones_like(x0['net']['graph']['A'])


def initialize_with_data(population, data, x0, Ns=None):
    initialize_with_dense_graph(population, data, x0)
