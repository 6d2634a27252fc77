"""Neuron-sharded MAP coordinate descent over several GPUs (interface of pyglm/inference/parallel_coord_descent.py).

The reference maps the per-neuron GLM fits over IPython.parallel engines that each hold the whole data set, gathers the
fitted `x['glms'][n]`, fits the network on the client and sums the engines' log-probabilities
(parallel_coord_descent.py:57-157; its calls into `prep_glm_inference` are stale, so the serial `coord_descent`
semantics are followed).  Here every torch.distributed rank (one process per GPU) owns the
contiguous block of postsynaptic neurons `neuron_shard(N, world, rank)`: it runs the lock-step L-BFGS on its own columns
of the engine only, the fitted parameter rows are all-gathered as ONE tensor collective (NCCL on the ranks' GPUs), and
the log posterior is an all-reduced scalar.
"""
import numpy as np

from ..utils.parallel_util import allgather_columns, allreduce_sum, collective_device, neuron_shard, world_rank
from .coord_descent import fit_glms_batched, fit_network
from .smart_init import initialize_with_data


def parallel_log_p(population, x, n_lo, n_hi, group=None):
    """log p(x, data) with the likelihood and the GLM priors of neurons [n_lo, n_hi) evaluated on this rank and summed
    over ranks (the reference sums the engines' values on the client, parallel_util.py:30,78)."""
    world, rank = world_rank(group)
    lp = 0.0
    if n_hi > n_lo:
        scale = population.glm.lkhd_scale.get_value()
        for data in population.data_sequences:
            lp += scale * float(np.sum(population.ll_grad(x, n_lo, n_hi, data=data, grad=False)))
        lp += sum(population.glm.log_prior(x['glms'][n]) for n in range(n_lo, n_hi))
    if rank == 0:                                            # shared terms are counted once
        lp += population.latent.log_p(x.get('latent', {})) + population.network.log_p(x['net'])
    return float(allreduce_sum([np.array([lp])], device=collective_device(population.device, group), group=group)[0][0])


def gather_glm_params(population, x, n_lo, n_hi, group=None):
    """Every rank's fitted rows of the dense parameter matrix -> the full state on all ranks (one all-gather)."""
    world, _ = world_rank(group)
    if world == 1:
        return x
    P_mine = np.stack([population.glm_param_vector(x['glms'][n]) for n in range(n_lo, n_hi)]) if n_hi > n_lo \
        else np.zeros((0, len(population.glm_param_vector(x['glms'][0]))))
    P = allgather_columns(P_mine, population.N, device=collective_device(population.device, group), group=group)
    population.set_dense_glm_params(x, P)
    return x


def parallel_coord_descent(population, x0=None, maxiter=50, atol=1e-5, group=None, verbose=False):
    """MAP estimate with the GLM fits partitioned by postsynaptic neuron.  `x0` must be identical on all ranks (when it
    is drawn here, every rank must have seeded np.random identically).  Every rank returns the same state."""
    world, rank = world_rank(group)
    N = population.model['N']
    n_lo, n_hi = neuron_shard(N, world, rank)
    if x0 is None:
        x0 = population.sample()
    initialize_with_data(population, population.data_sequences[-1], x0)
    x = x0
    lp_prev = parallel_log_p(population, x, n_lo, n_hi, group)
    if verbose and rank == 0:
        print("Initial LP=%.2f." % lp_prev)
    converged, it = False, 0
    while not converged and it < maxiter:
        it += 1
        if n_hi > n_lo:
            fit_glms_batched(population, x, n_lo=n_lo, n_hi=n_hi)
        gather_glm_params(population, x, n_lo, n_hi, group)
        fit_network(population, x)                           # deterministic and cheap: every rank does the same
        lp = parallel_log_p(population, x, n_lo, n_hi, group)
        if verbose and rank == 0:
            print("Iteration %d: LP=%.2f. Change in LP: %.2f" % (it, lp, lp - lp_prev))
        converged = np.abs(lp - lp_prev) < atol
        lp_prev = lp
    return x
