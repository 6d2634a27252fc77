"""world_size-2 gloo test of the N>1 host path: time-sharded partial sums all-reduced into the whole
(the reduction bench.py --gpus N performs over NCCL) and neuron-sharded columns all-gathered."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyglm_oracle as orc
from tests.helpers import make_problem, rel_err
from theano_pyglm_b200.utils.parallel_util import allgather_columns, allreduce_sum, neuron_shard, time_shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = make_problem(3000, 6, 5, network=True, seed=21)
        R = p['ibasis'].shape[0]
        S = p['S'].astype(np.float64)
        # ---- time sharding: filter with halo, local ll/grad (the oracle stands in for the GPU), all-reduce
        lo, hi, halo = time_shard(p['T'], world, rank, R)
        fS = orc.convolve_with_basis_direct(S[lo - halo:hi], p['ibasis'])[halo:]
        ll, gb, gw = orc.population_ll_grad(fS, S[lo:hi], p['dt'], p['bias'], p['w'], p['A'], p['W'], orc.NLIN_SOFTPLUS)
        ll, gb, gw = allreduce_sum([ll, gb, gw])
        # ---- neuron sharding: each rank evaluates its own columns, all-gather
        n_lo, n_hi = neuron_shard(p['N'], world, rank)
        fS_full = orc.convolve_with_basis_direct(S, p['ibasis'])
        ll_cols = np.array([orc.glm_ll(fS_full, S, p['dt'], n, p['bias'][n], p['w'][n], p['A'], p['W'], orc.NLIN_SOFTPLUS)
                            for n in range(n_lo, n_hi)])
        ll_gathered = allgather_columns(ll_cols, p['N'])
        if rank == 0:
            q.put((ll, gb, gw, ll_gathered))
    finally:
        dist.destroy_process_group()


def test_time_and_neuron_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    ll, gb, gw, ll_gathered = q.get()
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p = make_problem(3000, 6, 5, network=True, seed=21)
    S = p['S'].astype(np.float64)
    fS = orc.convolve_with_basis_direct(S, p['ibasis'])
    ll0, gb0, gw0 = orc.population_ll_grad(fS, S, p['dt'], p['bias'], p['w'], p['A'], p['W'], orc.NLIN_SOFTPLUS)
    assert rel_err(ll, ll0) < 1e-12 and rel_err(gb, gb0) < 1e-10 and rel_err(gw, gw0) < 1e-10
    assert rel_err(ll_gathered, ll0) < 1e-12


def _splice_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from theano_pyglm_b200.inference.parallel_gibbs import concatenate_parallel_updates
        from theano_pyglm_b200.models.model_factory import make_model, stabilize_sparsity
        from theano_pyglm_b200.population import Population
        N = 5
        model = make_model('sparse_weighted_model', N=N, dt=0.001)
        stabilize_sparsity(model)
        popn = Population(model)                                      # components only: no GPU is touched
        np.random.seed(4)
        x = popn.sample()                                             # the same start on both ranks
        n_lo, n_hi = neuron_shard(N, world, rank)
        W = x['net']['weights']['W'].reshape(N, N)
        for n in range(n_lo, n_hi):                                   # every rank rewrites its own columns only
            x['glms'][n]['bias']['bias'] = np.array([100.0 * rank + n])
            x['glms'][n]['imp']['g_%d' % n] = np.full(5, 1.0 + rank + 0.1 * n)
            x['net']['graph']['A'][:, n] = (np.arange(N) + n + rank) % 2
            W[:, n] = 10.0 * rank + n + 0.01 * np.arange(N)
        x['net']['weights']['W'] = W.ravel()
        x = concatenate_parallel_updates(popn, x, n_lo, n_hi)
        q.put((rank, x))
    finally:
        dist.destroy_process_group()


def test_neuron_sharded_state_splice_world2():
    """concatenate_parallel_updates (parallel_gibbs.py:24-37 as two tensor all-gathers): after the splice both ranks
    hold every owner's columns of A / W and every owner's GLM variables."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_splice_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = dict(q.get() for _ in range(world))
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    N = 5
    for rank in range(world):
        x = res[rank]
        W = x['net']['weights']['W'].reshape(N, N)
        assert x['net']['graph']['A'].dtype == np.int8
        for n in range(N):
            owner = 0 if n < neuron_shard(N, world, 0)[1] else 1
            assert x['glms'][n]['bias']['bias'][0] == 100.0 * owner + n
            assert np.array_equal(x['glms'][n]['imp']['g_%d' % n], np.full(5, 1.0 + owner + 0.1 * n))
            assert np.array_equal(x['net']['graph']['A'][:, n], (np.arange(N) + n + owner) % 2)
            assert np.array_equal(W[:, n], 10.0 * owner + n + 0.01 * np.arange(N))


class _OracleHandle:
    """Stands in for engine.Dataset in a CPU test of the host logic above it: evaluates this rank's time shard with the
    float64 oracle (the product itself never does this: it has no CPU path)."""

    def __init__(self, S, halo, dt, ibasis):
        self.fS = orc.convolve_with_basis_direct(S, ibasis)[halo:]
        self.S, self.dt = S[halo:], dt

    def ll_grad(self, bias, w, A, W, nlin=None, n_lo=0, n_hi=None, path=None, grad=True, w_stim=None):
        N = self.S.shape[1]
        n_hi = N if n_hi is None else n_hi
        A = np.ones((N, N), np.int8) if A is None else A             # null network = complete graph, unit weights
        W = np.ones((N, N)) if W is None else np.asarray(W).reshape(N, N)
        ll, gb, gw = orc.population_ll_grad(self.fS, self.S, self.dt, bias, w.reshape(N, N, -1), A, W, orc.NLIN_SOFTPLUS)
        if not grad:
            return ll[n_lo:n_hi]
        return ll[n_lo:n_hi], gb[n_lo:n_hi], gw.reshape(N, -1)[n_lo:n_hi]


def _tshard_pop_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from theano_pyglm_b200.models.model_factory import make_model
        from theano_pyglm_b200.population import Population
        from theano_pyglm_b200.utils.parallel_util import shard_data_by_time
        N = 4
        popn = Population(make_model('standard_glm', N=N, dt=0.001), time_sharded=True)
        rng = np.random.default_rng(8)
        S = (rng.random((2501, N)) < 0.03).astype(float)
        ib = popn.glm.imp_model.ibasis
        mine = shard_data_by_time({'S': S, 'N': N, 'dt': 0.001, 'T': 2.501}, ib.shape[0])
        mine['_b200'] = _OracleHandle(mine['S'], mine['halo'], 0.001, ib)
        mine['preprocessed'] = True
        popn.data_sequences.append(mine)
        popn.set_data(mine)
        np.random.seed(3)
        x = popn.sample()
        whole = {'S': S, '_b200': _OracleHandle(S, 0, 0.001, ib), 'preprocessed': True}
        ref = Population(make_model('standard_glm', N=N, dt=0.001))
        ref.data_sequences.append(whole)
        ref.set_data(whole)
        f, g = popn.glm_log_p_grad(x, 1)
        f0, g0 = ref.glm_log_p_grad(x, 1)
        q.put((rank, popn.compute_log_p(x), ref.compute_log_p(x), f, f0, g, g0))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_time_sharded_population_sums_over_ranks_world2():
    """Population(time_sharded=True) on two gloo ranks, each holding half of the recording (plus the filter's left context):
    log p and the per-neuron objective / gradient of coordinate descent equal those of the whole recording on one handle."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tshard_pop_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    for rank, lp, lp0, f, f0, g, g0 in res:
        assert abs(lp - lp0) < 1e-10 * abs(lp0)
        assert abs(f - f0) < 1e-10 * abs(f0) and np.max(np.abs(g - g0)) < 1e-9 * np.max(np.abs(g0))
    assert res[0][1] == res[1][1]
