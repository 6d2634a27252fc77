"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The reference parallelises over postsynaptic neurons with IPython.parallel: every engine holds the
whole data set, `dview.map_async` scatters neuron indices, the client splices the returned columns
and sums scalar log-probabilities (pyglm/utils/parallel_util.py:8-31,154-183;
inference/parallel_gibbs.py:24-37,162-168).  Here:

  * neuron sharding  -- rank g owns postsynaptic columns [n_lo, n_hi): `neuron_shard`; results are
    all-gathered (`allgather_columns`) and scalar log-p all-reduced (`allreduce_sum`);
  * time sharding    -- rank g owns bins [lo, hi) plus an R-bin left halo of spikes for the filter:
    `time_shard`; ll / gradient partial sums are all-reduced, the same algebra the reference uses
    to add data sequences (population.py:41-43, coord_descent.py:52-57).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world_rank(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def collective_device(device_index=0, group=None):
    """Where a collective's buffers must live: this rank's GPU under NCCL, host memory under gloo (CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", int(device_index))
    return torch.device("cpu")


def neuron_shard(N, world_size, rank):
    """Contiguous block of postsynaptic neurons for `rank` (sizes differ by at most one)."""
    base, extra = divmod(N, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def time_shard(T, world_size, rank, R):
    """(lo, hi, halo): bins [lo, hi) are this rank's; it also needs bins [lo-halo, lo) as filter context."""
    lo, hi = neuron_shard(T, world_size, rank)
    return lo, hi, min(R, lo)


def shard_data_by_time(data, R, group=None):
    """This rank's share of a recording for a time-sharded Population: bins [lo, hi) plus the R-bin left context the
    spike-history filter needs (`halo`), as a data dictionary ready for `Population.add_data`."""
    world, rank = world_rank(group)
    T = data['S'].shape[0]
    lo, hi, halo = time_shard(T, world, rank, R)
    out = dict(data)
    out['S'] = np.ascontiguousarray(data['S'][lo - halo:hi])
    out['halo'] = halo
    out.pop('_b200', None)
    out['preprocessed'] = False
    return out


def _as_tensor(a, device):
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def allreduce_sum(arrays, device="cpu", group=None):
    """Sum a list of float64 arrays over ranks with ONE collective (they are packed into one buffer).
    Returns numpy arrays shaped like the inputs."""
    arrays = [np.asarray(a, dtype=np.float64) for a in arrays]
    flat = np.concatenate([a.ravel() for a in arrays]) if arrays else np.zeros(0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        buf = _as_tensor(flat, device)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        flat = buf.cpu().numpy()
    out, off = [], 0
    for a in arrays:
        out.append(flat[off:off + a.size].reshape(a.shape))
        off += a.size
    return out


def allgather_columns(local, N, device="cpu", group=None):
    """Gather per-neuron results computed on a neuron shard into the full population order.

    `local` has shape (n_hi - n_lo, ...) for this rank's shard; returns (N, ...).  This is the
    splice the reference does on the client (parallel_gibbs.py:24-37) as one all_gather."""
    local = np.ascontiguousarray(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == N
        return local
    world = dist.get_world_size(group)
    widths = [np.subtract(*neuron_shard(N, world, r)[::-1]) for r in range(world)]
    wmax = max(widths)
    pad = np.zeros((wmax,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    mine = _as_tensor(pad, device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p.cpu().numpy()[:w] for p, w in zip(parts, widths)], axis=0)


def make_peer_comm(max_doubles, device, group=None):
    """engine.PeerComm for the current torch.distributed job: the 128-byte cudaIpc handles are exchanged
    with all_gather_object (any backend), after which the all-reduce runs over NVLink peer memory with no
    further use of torch.distributed."""
    from ..engine import PeerComm
    if not (dist.is_available() and dist.is_initialized()):
        return PeerComm(0, 1, device, max_doubles, lambda b: [b])
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def exchange(mine):
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
        return parts
    return PeerComm(rank, world, device, max_doubles, exchange)


def splice_network_columns(A, W, n_lo, n_hi, A_cols, W_cols):
    """Write back the columns a rank resampled: A[:, n], W[:, n] for n in [n_lo, n_hi)."""
    A[:, n_lo:n_hi] = A_cols
    W[:, n_lo:n_hi] = W_cols
    return A, W
