"""Neuron-sharded MCMC over several GPUs (interface of pyglm/inference/parallel_gibbs.py).

The reference scatters neuron indices over IPython.parallel engines, every engine resamples the variables
of its neurons against a full copy of the data, and the client splices the results
(`concatenate_parallel_updates`, parallel_gibbs.py:24-37; `parallel_gibbs_sample`, :40-197).  Here every
torch.distributed rank (one process per GPU) holds the data set, owns the contiguous block of postsynaptic
neurons `neuron_shard(N, world, rank)`, runs the HMC updates and the lock-step collapsed Gibbs sweep for its
own columns only, and the splice is one all-gather per iteration.  Columns are conditionally independent
given A, W and the data (population.py:28-32), so no collective is needed inside a sweep.
"""
import copy

import numpy as np
import torch.distributed as dist

from ..utils.parallel_util import allgather_columns, collective_device, neuron_shard, world_rank
from .gibbs import initial_state, initialize_batched_updates, initialize_updates
from .parallel_coord_descent import parallel_log_p


def _world(group=None):
    return world_rank(group)


def concatenate_parallel_updates(population, x, n_lo, n_hi, group=None):
    """Splice the columns every rank resampled into one consistent state on all ranks (parallel_gibbs.py:24-37):
    x['glms'][n], A[:, n] and W[:, n] come from the owner of neuron n.  Two tensor collectives on the ranks' own
    devices (NCCL on GPUs): the float64 rows [GLM parameter vector | W[:, n]] and the int8 rows A[:, n] -- no pickling,
    no object gather."""
    world, rank = _world(group)
    if world == 1:
        return x
    N = population.N
    dev = collective_device(population.device, group)
    W = x['net']['weights']['W'].reshape(N, N)
    P = population.dense_glm_params(x)
    D = P.shape[1]
    mine = np.concatenate([P[n_lo:n_hi], W[:, n_lo:n_hi].T], axis=1)                   # (n, D + N) float64
    full = allgather_columns(np.ascontiguousarray(mine), N, device=dev, group=group)
    A_cols = allgather_columns(np.ascontiguousarray(x['net']['graph']['A'][:, n_lo:n_hi].T), N, device=dev, group=group)
    population.set_dense_glm_params(x, full[:, :D])
    x['net']['graph']['A'] = np.ascontiguousarray(A_cols.T).astype(np.int8)
    x['net']['weights']['W'] = np.ascontiguousarray(full[:, D:].T).ravel()
    return x


def parallel_gibbs_sample(population, N_samples=1000, x0=None, init_from_mle=False, callback=None, seed=None,
                          group=None, verbose=False):
    """parallel_gibbs.py:40-197 over torch.distributed.  Every rank returns the same list of samples.
    `x0` must be identical on all ranks (rank 0's state is broadcast when it is drawn here)."""
    world, rank = _world(group)
    N = population.model['N']
    n_lo, n_hi = neuron_shard(N, world, rank)
    if x0 is None:
        box = [initial_state(population, init_from_mle, verbose) if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        x0 = box[0]
    if seed is not None:
        np.random.seed(seed + 7919 * rank)                      # independent streams per shard
    serial_updates, parallel_updates = initialize_updates(population)
    batched_updates = initialize_batched_updates(population)
    net_update = parallel_updates[-1]
    x = x0
    x_smpls = [copy.deepcopy(x0)]
    for smpl in range(N_samples):
        if callback is not None and rank == 0:
            callback(x)
        if verbose:                                               # a collective: every rank takes part
            lp = parallel_log_p(population, x, n_lo, n_hi, group)
            if rank == 0:
                print("Gibbs iteration %d. Log prob: %.3f" % (smpl, lp))
        if n_hi > n_lo:
            for upd in batched_updates:                           # this rank's neurons, in lock-step
                upd.update(x, n_lo, n_hi)
        if n_hi > n_lo:
            net_update.begin(x, n_lo, n_hi)
            net_update.sweep_batched(x, n_lo, n_hi)
            net_update.end()
        x = concatenate_parallel_updates(population, x, n_lo, n_hi, group=group)
        for upd in serial_updates:                                # none for the supported models
            upd.update(x)
        x_smpls.append(copy.deepcopy(x))
    return x_smpls
