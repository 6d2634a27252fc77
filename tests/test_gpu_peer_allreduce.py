"""GPU test of the peer-memory all-reduce (csrc/allreduce.cu) and of time-sharded ll+gradient through it.

The ranks are separate processes, as in production, but share cuda:0 (cudaIpc works between processes on one
device), so the test runs on a single-GPU box; the handles travel over a gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyglm_oracle as orc
from tests.helpers import make_problem, rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import theano_pyglm_b200 as pg
        from theano_pyglm_b200.utils.parallel_util import make_peer_comm, time_shard
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        stream = torch.cuda.current_stream()
        comm = make_peer_comm(400000, 0)
        # 1. plain vectors, several sizes and many back-to-back calls (epoch / double-buffer logic), in place
        ok = True
        for n in (1, 37, 3672, 300001):
            for it in range(6):
                x = torch.arange(n, dtype=torch.float64, device=dev) * (rank + 1) + it
                comm.allreduce_sum_dev(x.data_ptr(), x.data_ptr(), n, stream.cuda_stream)
                ref = torch.arange(n, dtype=torch.float64, device=dev) * sum(r + 1 for r in range(world)) + it * world
                ok &= bool(torch.equal(x, ref))
        # 2. time-sharded ll + gradient: K1 with an R-bin halo, tensor-core path, partials summed over ranks
        p = make_problem(6000, 6, 5, network=True, seed=21)
        R = p['ibasis'].shape[0]
        lo, hi, halo = time_shard(p['T'], world, rank, R)
        ds = pg.Dataset(p['S'][lo - halo:hi], p['dt'], p['ibasis'], halo=halo)
        N, NB = p['N'], p['N'] * p['B']
        d_bias = torch.from_numpy(p['bias']).to(dev)
        d_w = torch.from_numpy(p['w'].reshape(N, NB)).to(dev)
        d_A = torch.from_numpy(p['A']).to(dev)
        d_W = torch.from_numpy(p['W']).to(dev)
        out = torch.zeros(N * (2 + NB), dtype=torch.float64, device=dev)
        ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), d_A.data_ptr(), d_W.data_ptr(), "explinear", 0, N, "auto",
                       out[:N].data_ptr(), out[N:2 * N].data_ptr(), out[2 * N:].data_ptr(), stream.cuda_stream)
        comm.allreduce_sum_dev(out.data_ptr(), out.data_ptr(), out.numel(), stream.cuda_stream)
        torch.cuda.synchronize()
        # 3. the same through the one-call form (here it falls back to evaluation + collective: 30 features)
        out1 = torch.zeros_like(out)
        ds.ll_grad_allreduce_dev(comm, d_bias.data_ptr(), d_w.data_ptr(), d_A.data_ptr(), d_W.data_ptr(), "explinear", "auto",
                                 out1.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ok &= bool(torch.equal(out1, out))
        # 4. a population the fused kernel covers in one launch (N = 27, 135 features): its final reduction carries the sum
        #    over ranks; bitwise equal to evaluation + collective, also when the two forms and plain all-reduces alternate
        p2 = make_problem(5000, 27, 5, network=True, seed=5)
        lo2, hi2, halo2 = time_shard(p2['T'], world, rank, R)
        ds2 = pg.Dataset(p2['S'][lo2 - halo2:hi2], p2['dt'], p2['ibasis'], halo=halo2)
        N2, NB2 = 27, 135
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        b2, w2, A2, W2 = t(p2['bias']), t(p2['w'].reshape(N2, NB2)), t(p2['A']), t(p2['W'])
        ref2 = torch.zeros(N2 * (2 + NB2), dtype=torch.float64, device=dev)
        ds2.ll_grad_dev(b2.data_ptr(), w2.data_ptr(), A2.data_ptr(), W2.data_ptr(), "explinear", 0, N2, "tc",
                        ref2[:N2].data_ptr(), ref2[N2:2 * N2].data_ptr(), ref2[2 * N2:].data_ptr(), stream.cuda_stream)
        comm.allreduce_sum_dev(ref2.data_ptr(), ref2.data_ptr(), ref2.numel(), stream.cuda_stream)
        for it in range(5):
            got2 = torch.full_like(ref2, float("nan"))
            ds2.ll_grad_allreduce_dev(comm, b2.data_ptr(), w2.data_ptr(), A2.data_ptr(), W2.data_ptr(), "explinear", "tc",
                                      got2.data_ptr(), stream.cuda_stream)
            if it % 2:
                x = torch.ones(1000, dtype=torch.float64, device=dev)
                comm.allreduce_sum_dev(x.data_ptr(), x.data_ptr(), 1000, stream.cuda_stream)
            torch.cuda.synchronize()
            ok &= bool(torch.equal(got2, ref2))
        ds2.close()
        q.put((rank, ok, out.cpu().numpy()))
        dist.barrier()
        comm.close()
        ds.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_allreduce_and_time_sharded_ll_grad(world, engine_lib):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = {}
    for _ in range(world):
        rank, ok, out = q.get(timeout=240)
        results[rank] = (ok, out)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p = make_problem(6000, 6, 5, network=True, seed=21)
    N = p['N']
    S = p['S'].astype(np.float64)
    fS = orc.convolve_with_basis_direct(S, p['ibasis'])
    ll0, gb0, gw0 = orc.population_ll_grad(fS, S, p['dt'], p['bias'], p['w'], p['A'], p['W'], orc.NLIN_SOFTPLUS)
    for rank in range(world):
        ok, out = results[rank]
        assert ok, "rank %d: all-reduce of plain vectors is wrong" % rank
        assert np.array_equal(out, results[0][1])                    # same summation order: bitwise identical
        assert rel_err(out[:N], ll0) < 1e-6 and rel_err(out[N:2 * N], gb0) < 1e-5
        assert rel_err(out[2 * N:], gw0.reshape(-1)) < 1e-5


def _gibbs_worker(rank, world, port, q, x_dtype=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from theano_pyglm_b200.inference.parallel_gibbs import parallel_gibbs_sample
        from theano_pyglm_b200.models.model_factory import make_model, stabilize_sparsity
        from theano_pyglm_b200.population import Population
        N = 5
        model = make_model('sparse_weighted_model', N=N, dt=0.001)
        stabilize_sparsity(model)
        popn = Population(model, x_dtype=x_dtype)
        rng = np.random.default_rng(3)                               # same data on every rank
        S = (rng.random((6000, N)) < 0.03).astype(float)
        popn.add_data({'S': S, 'N': N, 'dt': 0.001, 'T': 6.0, 'stim': None, 'dt_stim': 0.1})
        np.random.seed(1)
        x0 = popn.sample()                                           # identical start on every rank
        smpls = parallel_gibbs_sample(popn, N_samples=2, x0=x0, seed=100)
        lp = popn.compute_log_p(smpls[-1])
        q.put((rank, smpls[0], smpls[-1], lp))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("x_dtype", [None, "none"])
def test_neuron_sharded_gibbs_splices_to_one_state(engine_lib, x_dtype):
    """parallel_gibbs_sample (parallel_gibbs.py:40-197): two ranks each resample their own columns on the
    engine; after the all-gather splice both hold the same state, and every column was touched.  x_dtype="none": the
    ranks hold the spike trains only -- the HMC updates evaluate ll / gradients from the spikes (from-spikes K2) and the
    collapsed A/W updates gather their currents from them (from-spikes K4)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gibbs_worker, args=(r, world, port, q, x_dtype)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, first, last, lp = q.get(timeout=300)
        res[rank] = (first, last, lp)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    (f0, l0, lp0), (f1, l1, lp1) = res[0], res[1]
    N = 5
    assert np.isfinite(lp0) and lp0 == lp1
    assert l0['net']['graph']['A'].dtype == np.int8
    assert np.array_equal(l0['net']['graph']['A'], l1['net']['graph']['A'])
    assert np.array_equal(l0['net']['weights']['W'], l1['net']['weights']['W'])
    W0 = f0['net']['weights']['W'].reshape(N, N)
    Wl = l0['net']['weights']['W'].reshape(N, N)
    for n in range(N):
        assert l0['glms'][n]['bias']['bias'] == l1['glms'][n]['bias']['bias']
        assert l0['glms'][n]['n'] == n
        assert not np.array_equal(W0[:, n], Wl[:, n])                # columns of both shards were resampled
        assert np.all(np.diag(l0['net']['graph']['A']) == 1)         # self edges stay (p_A = 1 - 1e-8 on the diagonal)


def _map_worker(rank, world, port, q, x_dtype=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from theano_pyglm_b200.inference.coord_descent import coord_descent
        from theano_pyglm_b200.inference.parallel_coord_descent import parallel_coord_descent, parallel_log_p
        from theano_pyglm_b200.models.model_factory import make_model
        from theano_pyglm_b200.population import Population
        from theano_pyglm_b200.utils.parallel_util import neuron_shard
        N = 5
        model = make_model('standard_glm', N=N, dt=0.001)
        popn = Population(model, x_dtype=x_dtype)
        rng = np.random.default_rng(11)                              # same data on every rank
        S = (rng.random((8000, N)) < 0.03).astype(float)
        popn.add_data({'S': S, 'N': N, 'dt': 0.001, 'T': 8.0, 'stim': None, 'dt_stim': 0.1})
        np.random.seed(2)
        x0 = popn.sample()
        for n in range(N):
            x0['glms'][n]['imp']['w_ir'] *= 0.01
        import copy
        x_par = parallel_coord_descent(popn, x0=copy.deepcopy(x0), maxiter=3)
        n_lo, n_hi = neuron_shard(N, world, rank)
        lp_par = parallel_log_p(popn, x_par, n_lo, n_hi)             # a collective
        x_ser = coord_descent(popn, x0=copy.deepcopy(x0), maxiter=3) if rank == 0 else None
        q.put((rank, popn.dense_glm_params(x_par), lp_par, popn.compute_log_p(x_par),
               None if x_ser is None else popn.dense_glm_params(x_ser),
               None if x_ser is None else popn.compute_log_p(x_ser)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("x_dtype", [None, "none"])
def test_neuron_sharded_map_matches_the_serial_coordinate_descent(engine_lib, x_dtype):
    """parallel_coord_descent (parallel_coord_descent.py:57-157): two ranks fit their own neurons' GLMs on the engine,
    all-gather the fitted rows and all-reduce the log posterior; the result is the serial coord_descent's, because the
    per-neuron problems are independent given the (constant) network.  x_dtype="none": every rank holds the spike trains
    only and evaluates its neurons from the spikes (from-spikes K2) -- the reference's layout, each engine with the full data."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_map_worker, args=(r, world, port, q, x_dtype)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, P, lp_par, lp_full, P_ser, lp_ser = q.get(timeout=300)
        res[rank] = (P, lp_par, lp_full, P_ser, lp_ser)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    assert np.array_equal(res[0][0], res[1][0])                     # one state on both ranks
    assert res[0][1] == res[1][1]                                   # the all-reduced log posterior
    assert abs(res[0][1] - res[0][2]) < 1e-7 * abs(res[0][2])       # ... equals the single-process evaluation
    # ... and the serial driver's optimum: the same log posterior to 1e-6; the parameters agree as far as the posterior
    # determines them (a shard's engine call and the full call differ at ~1e-9, so the two L-BFGS runs stop at slightly
    # different points of a flat optimum: measured 0.02 on coefficients of size 1..20)
    assert abs(res[0][2] - res[0][4]) < 1e-6 * abs(res[0][4]), (res[0][2], res[0][4])
    d = np.abs(res[0][0] - res[0][3])
    assert np.max(d) < 0.1, float(np.max(d))


def _tshard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import copy
        from theano_pyglm_b200.inference.coord_descent import coord_descent
        from theano_pyglm_b200.models.model_factory import make_model
        from theano_pyglm_b200.population import Population
        from theano_pyglm_b200.utils.parallel_util import shard_data_by_time
        N = 5
        model = make_model('standard_glm', N=N, dt=0.001)
        rng = np.random.default_rng(12)                              # the same recording on every rank
        S = (rng.random((9001, N)) < 0.03).astype(float)
        data = {'S': S, 'N': N, 'dt': 0.001, 'T': 9.001, 'stim': None, 'dt_stim': 0.1}
        popn = Population(model, time_sharded=True)
        R = popn.glm.imp_model.ibasis.shape[0]
        popn.add_data(shard_data_by_time(data, R))
        np.random.seed(4)
        x0 = popn.sample()
        for n in range(N):
            x0['glms'][n]['imp']['w_ir'] *= 0.01
        lp0 = popn.compute_log_p(x0)
        f0, g0 = popn.glm_log_p_grad(x0, 2)
        x_map = coord_descent(popn, x0=copy.deepcopy(x0), maxiter=2)
        res = dict(lp0=lp0, f0=f0, g0=g0, P=popn.dense_glm_params(x_map), lp=popn.compute_log_p(x_map))
        if rank == 0:                                                # the whole recording on one handle
            full = Population(model)
            full.add_data(dict(data))
            fs, gs = full.glm_log_p_grad(x0, 2)
            xs = coord_descent(full, x0=copy.deepcopy(x0), maxiter=2)
            res.update(lp0_full=full.compute_log_p(x0), f0_full=fs, g0_full=gs, P_full=full.dense_glm_params(xs),
                       lp_full=full.compute_log_p(xs))
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_time_sharded_population_equals_the_whole_recording(engine_lib):
    """Population(time_sharded=True): every rank holds its time shard (with the filter's left context) and every ll /
    gradient is summed over the ranks, so log p, the per-neuron objective of coordinate descent and the MAP estimate are
    those of the whole recording, identical on all ranks -- the serial drivers run unchanged."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tshard_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        rank, r = q.get(timeout=300)
        res[rank] = r
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    a, b = res[0], res[1]
    assert a['lp0'] == b['lp0'] and a['lp'] == b['lp'] and np.array_equal(a['P'], b['P'])      # one state on every rank
    assert abs(a['lp0'] - a['lp0_full']) < 1e-7 * abs(a['lp0_full'])
    assert abs(a['f0'] - a['f0_full']) < 1e-7 * abs(a['f0_full'])
    assert np.max(np.abs(a['g0'] - a['g0_full'])) < 1e-5 * np.max(np.abs(a['g0_full']))
    assert abs(a['lp'] - a['lp_full']) < 1e-6 * abs(a['lp_full'])
    assert np.max(np.abs(a['P'] - a['P_full'])) < 0.1
