#!/usr/bin/env python
"""Benchmark of the pyglm hot path on B200: population GLM ll+grad evals/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One "step" = one ll+gradient evaluation for all N neurons of one data sequence
(BASELINE.json metric).  Workload at one GPU: config C2
(standard_glm, N=27, T=10^6 bins, B=5, R=200), synthetic Bernoulli(0.02) spikes.
With --gpus G (torchrun, one rank per GPU) every rank owns its own T=10^6-bin sequence
(time-sharded, weak scaling) and the per-step partial sums are added over ranks -- by the engine's
one-shot all-reduce over NVLink peer memory (csrc/allreduce.cu; --allreduce nccl uses NCCL instead) --
exactly as the reference sums ll over data sequences (coord_descent.py:52-57).

Prints ONE JSON line on rank 0 (contract in the task statement).  At one GPU the line also carries an `extra`
array -- the other half of BASELINE.json's metric and the larger configurations, each with its own roofline, e2e,
clocks and cpu_baseline: the K1 filter, C3 ll+grad (tensor-bound GEMM path), C3 Gibbs edge-sweeps/s, and one of eight
time shards of C4 (`--no-extra` skips them).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, T, B, network, description)
    "c1": dict(N=4, T=60_000, B=5, desc="standard_glm N=4 T=60s (C1)"),
    "c2": dict(N=27, T=1_000_000, B=5, desc="standard_glm N=27 T=1e6 bins B=5 R=200 (C2)"),
    "c3": dict(N=256, T=1_000_000, B=5, desc="network GLM N=256 T=1e6 bins B=5 (C3, ll+grad part)"),
    # one GPU's share of C4 (N=1024, T=4e6, B=10) under 8-way time sharding; planes-only ingest (no FP32 X resident)
    "c4-shard": dict(N=1024, T=500_000, B=10, x_dtype="planes",
                     desc="network-size GLM N=1024 B=10, T=5e5 bins = one of 8 time shards of C4 (ll+grad)"),
    # second headline metric: collapsed Gibbs over A/W, batched delta-ll (sparse_weighted_model)
    "c3-gibbs": dict(N=256, T=1_000_000, B=5, gibbs=True, desc="network GLM N=256 T=1e6 bins: collapsed Gibbs over A/W (C3)"),
    "gibbs-small": dict(N=32, T=200_000, B=5, gibbs=True, desc="network GLM N=32 T=2e5 bins: collapsed Gibbs over A/W (smoke size)"),
    # C5's population (N=4096, B=5) time-sharded: every GPU owns 2^17 bins (10.7 GB of planes; C5 proper is 1.25e6 bins per
    # GPU = 102 GB), planes-only ingest; the per-step collective is the all-reduce of ll / g_bias / G = 671 MB
    "c5-shard": dict(N=4096, T=131_072, B=5, x_dtype="planes",
                     desc="N=4096 B=5, T=2^17 bins per GPU: time shard of the C5 population (ll+grad, 671 MB all-reduce per step)"),
}
# the same population with B = 10 basis functions ("C5 ... B=5 default; also report B=10"): 2^16 bins per GPU
# (10.7 GB of planes), all-reduce of 1.34 GB per step
WORKLOADS["c5-shard-b10"] = dict(N=4096, T=65_536, B=10, x_dtype="planes",
                                 desc="N=4096 B=10, T=2^16 bins per GPU: time shard of the C5 population with B=10 (ll+grad, 1.34 GB all-reduce per step)")
# from-spikes K2: spikes-only datasets, operand planes produced by K1 inside every evaluation (never resident)
WORKLOADS["c2-from-spikes"] = dict(WORKLOADS["c2"], x_dtype="none", desc="C2 evaluated from the spike trains (no X resident): standard_glm N=27 T=1e6 bins B=5")
WORKLOADS["c4-shard-from-spikes"] = dict(WORKLOADS["c4-shard"], x_dtype="none",
                                         desc="the C4 time shard (N=1024 B=10, T=5e5 bins) evaluated from the spike trains: 0.5 GB of spikes "
                                              "and 2 GB of chunk buffers resident instead of 20.5 GB of planes")
WORKLOADS["c4-neuron-shard"] = dict(N=1024, T=250_000, B=10, x_dtype="none",
                                    desc="N=1024 B=10, T=2.5e5 bins, all presynaptic spike trains on the GPU, its share of the postsynaptic "
                                         "neurons evaluated from the spikes (C4's split by postsynaptic neuron; C4 proper is T=4e6)")
METRIC = "GLM ll+grad evals/sec"
UNIT = "evals/s"


def make_inputs(wl, seed):
    """Synthetic C2-style inputs: Bernoulli(0.02) spikes, raised-cosine basis,
    bias ~ N(20, 0.1^2), w_ir ~ N(0, 0.05^2), complete graph, unit weights, softplus."""
    from theano_pyglm_b200.utils.basis import make_standard_ibasis
    N, T, B = wl["N"], wl["T"], wl["B"]
    rng = np.random.default_rng(seed)
    if T * N <= 300_000_000:
        S = (rng.random((T, N)) < 0.02).astype(np.uint8)
    else:                                   # large shards: one random byte per bin, p = 5/256
        S = (rng.integers(0, 256, size=(T, N), dtype=np.uint8) < 5).astype(np.uint8)
    ib = make_standard_ibasis(B=B, dt=0.001, dt_max=0.2)
    bias = 20.0 + 0.1 * rng.standard_normal(N)
    w = 0.05 * rng.standard_normal((N, N * B))
    return dict(S=S, ibasis=ib, bias=bias, w=w, dt=0.001)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
        return float(pk["bf16_tflops"]), "measured burst bf16 (MEASURED_PEAKS.json)"
    return 1590.0, "fallback (B200_PROFILING.md)"


def workload_config(wl):
    """The `config` object of the JSON line: the workload and nothing measured, identical in both arms."""
    return {"workload": wl["desc"], "N": wl["N"], "T_bins_per_gpu": wl["T"], "B": wl["B"], "R": 200, "nlin": "explinear",
            "l2": "inputs_exceed_l2"}


def median_block(times):
    return float(np.median(np.asarray(times, dtype=np.float64)))


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores (bench.py may execute oracle/ only here)
# --------------------------------------------------------------------------------------
def cpu_prepare(inp, wl, T_sample, shape="gemm"):
    """Returns a closure running one ll+grad eval of all N neurons on the first T_sample bins with the
    float64 oracle.  The filtered spike train is prepared outside (it is resident data for both arms)."""
    from oracle import pyglm_oracle as orc
    N, B = wl["N"], wl["B"]
    S = inp["S"][:T_sample].astype(np.float64)
    fS = orc.convolve_with_basis_direct(S, inp["ibasis"])
    w3 = inp["w"].reshape(N, N, B)
    A = np.ones((N, N), dtype=np.int8)
    W = np.ones((N, N))
    if shape == "gemm":
        return lambda: orc.population_ll_grad(fS, S, inp["dt"], inp["bias"], w3, A, W, orc.NLIN_SOFTPLUS)

    def per_neuron():   # reference-shaped: python loop over neurons, T x N x B product materialised (impulse.py:58)
        for n in range(N):
            orc.glm_ll_grad(fS, S, inp["dt"], n, inp["bias"][n], w3[n], A, W, orc.NLIN_SOFTPLUS)
    return per_neuron


def cpu_eval_time(inp, wl, T_sample, budget_s, shape="gemm"):
    """Mean seconds for one CPU eval on the first T_sample bins: one warm-up call, then repeats until `budget_s`
    seconds of timed work (at least 2).  Returns (seconds per eval, repeats)."""
    fn = cpu_prepare(inp, wl, T_sample, shape)
    fn()
    n, t0 = 0, time.perf_counter()
    while n < 2 or time.perf_counter() - t0 < budget_s:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n, n


def run_reference(args, wl):
    """The CPU arm: the float64 oracle port of the reference's path, whole-population BLAS form, ALL bins of the
    workload every step (no extrapolation), all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = make_inputs(wl, seed=1234)
    T_sample = wl["T"] if wl["N"] <= 64 else min(wl["T"], 50_000)
    scale = wl["T"] / T_sample
    cores = os.cpu_count() or 1
    fn = cpu_prepare(inp, wl, T_sample, "gemm")
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt_step = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (dt_step * scale)
    sample = ("every step evaluates all %d bins" % wl["T"] if scale == 1 else
              "first %d of %d bins per step, evals/s scaled by %d/%d" % (T_sample, wl["T"], T_sample, wl["T"])) + \
             "; float64 oracle port, whole-population BLAS GEMM form (filter excluded: resident data in both arms)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_step * scale * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# Our arm
# --------------------------------------------------------------------------------------
class Timer:
    """Device timing of K-step blocks: barrier + synchronize on both sides, CUDA events on the launching stream, max
    over ranks.  `blocks()` repeats the K-step block until the timed total reaches `min_total_s` and reports the
    median block (each block is exactly K steps)."""

    def __init__(self, torch, dist, world, dev, stream):
        self.torch, self.dist, self.world, self.dev, self.stream = torch, dist, world, dev, stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def block(self, fn, steps):
        torch = self.torch
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            fn()
        ev1.record(self.stream)
        self.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def blocks(self, fn, steps, min_total_s=0.5, max_blocks=400):
        times = [self.block(fn, steps)]
        n = int(min(max_blocks, max(1, np.ceil(min_total_s * 1e3 / max(times[0], 1e-3)))))
        if self.world > 1:                   # every rank must run the same number of blocks
            nt = self.torch.tensor([n], device=self.dev)
            self.dist.all_reduce(nt, op=self.dist.ReduceOp.MAX)
            n = int(nt.item())
        for _ in range(n - 1):
            times.append(self.block(fn, steps))
        return median_block(times), times


def wall_blocks(fn, steps, sync, min_total_s=0.5, max_blocks=400):
    """Wall-clock timing of K-step blocks of a call that synchronises itself; median block in ms."""
    times = []
    while len(times) < max_blocks and (not times or sum(times) < min_total_s * 1e3):
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        sync()
        times.append((time.perf_counter() - t0) * 1e3)
    return median_block(times), times


def _traffic(kernel, workload_key):
    """dram__bytes_read + dram__bytes_write per launch of `kernel`, from the committed ncu capture (profiles/)."""
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as f:
        return json.load(f).get(kernel, {}).get(workload_key)


def llgrad_roofline(ds, wl, ms_kernel, path, workload_key):
    """roofline object of the ll+grad launch sequence: algorithmic bytes (HBM-bound fused kernel) or flops (GEMM path)."""
    N, T, B = wl["N"], wl["T"], wl["B"]
    NB = N * B
    peak, peak_src = load_peaks()
    info = ds.path_info(path)
    x_bytes = T * ds.ldx * 4
    passes = info.get("x_passes", 2)
    alg_bytes = passes * x_bytes + T * N + 16 * N * NB          # X once per contraction pass + S once + params/outputs
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    traffic = None          # dram__bytes_read+write of the dominant kernel, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(info.get("kernel", ""), {}).get(workload_key)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": info.get("kernel"), "peak_source": peak_src, "kernel_ms": ms_kernel,
            "algorithmic_bytes": alg_bytes}
    if info.get("bound") == "tensor":
        # algorithmic flops of one eval: two contractions, 2 flop per MAC; the split executes 3x as many
        tpeak, tsrc = load_tensor_peak()
        alg_flops = 4.0 * T * N * N * B
        tf = alg_flops / (ms_kernel * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                "traffic": traffic, "kernel": info.get("kernel"), "peak_source": tsrc, "kernel_ms": ms_kernel,
                "algorithmic_flops": alg_flops, "executed_over_algorithmic": 3.0,
                "note": "each contraction runs as 3 FP16 products of the error-free operand splits"}
    return roof, info


def llgrad_cpu_baseline(inp, wl, budget_s):
    """The float64 oracle on the host cores, bounded sample; BLAS-GEMM form plus the reference-shaped per-neuron loop."""
    N, T = wl["N"], wl["T"]
    T_s = min(T, 500_000 if N <= 64 else (50_000 if N <= 256 else 8_000))
    t_gemm, n_gemm = cpu_eval_time(inp, wl, T_s, budget_s, "gemm")
    t_gemm *= T / T_s
    T_r = min(T, 100_000 if N <= 64 else (10_000 if N <= 256 else 2_000))
    t_ref, n_ref = cpu_eval_time(inp, wl, T_r, budget_s / 2, "reference") if N <= 256 else (float("nan"), 0)
    t_ref *= T / T_r
    return {"value": 1.0 / t_gemm, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "float64 oracle, whole-population BLAS GEMM form: mean of %d evals on the first %d of %d bins "
                      "(~%d s of CPU work), scaled to the full recording; reference-shaped per-neuron loop "
                      "(impulse.py:58 style, %d evals on %d bins, scaled): %.4g evals/s"
                      % (n_gemm, T_s, T, budget_s, n_ref, T_r, 1.0 / t_ref),
            "reference_shaped_value": None if n_ref == 0 else 1.0 / t_ref}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import theano_pyglm_b200 as pg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    collective = "none (single GPU)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        collective = "nccl all_reduce"
        if args.allreduce == "p2p":
            # one-shot all-reduce over NVLink peer memory (csrc/allreduce.cu); every rank must agree on the choice
            from theano_pyglm_b200.utils.parallel_util import make_peer_comm
            ok = torch.ones(1, device=dev)
            try:
                comm = make_peer_comm(wl["N"] * (2 + wl["N"] * wl["B"]), local_rank)
            except Exception as exc:                       # no peer access on this box: NCCL does the sum
                sys.stderr.write("rank %d: peer all-reduce unavailable (%s); using NCCL\n" % (rank, exc))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1:
                comm = None
            else:
                collective = "sum over ranks folded into the evaluation's final reduction: one-shot exchange over NVLink peer memory (tc_final2_allreduce_kernel)"

    stream = torch.cuda.current_stream()
    timer = Timer(torch, dist, world, dev, stream)
    rec = llgrad_record(args, wl, args.workload, pg, torch, dist, timer, world, rank, local_rank, comm, collective,
                        min_total_s=0.5, cpu_budget_s=0.0 if args.no_cpu else 10.0)
    extra = None
    if not args.no_extra and args.workload == "c2":
        if world == 1:
            extra = run_extras(args, pg, torch, dist, timer, local_rank) if rank == 0 else None
        else:
            extra = run_scale_extras(args, pg, torch, dist, timer, world, rank, local_rank)     # collective: every rank
    if rank == 0:
        line = rec
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def llgrad_record(args, wl, workload_key, pg, torch, dist, timer, world, rank, local_rank, comm, collective,
                  min_total_s, cpu_budget_s):
    """One ll+grad workload measured three ways: device-resident steps (`value`), the reference-facing host-buffer call
    (`e2e`), and the launch sequence alone (roofline).  Returns the JSON record on rank 0 (None elsewhere)."""
    dev = timer.dev
    stream = timer.stream
    N, T, B = wl["N"], wl["T"], wl["B"]
    NB = N * B
    inp = make_inputs(wl, seed=1234 + rank)      # every rank owns its own sequence (time shard)
    t_ing0 = time.perf_counter()
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], device=local_rank, x_dtype=wl.get("x_dtype", "f32"))
    path = args.path
    nlin = "explinear"

    # resident parameters + outputs (device-timed path)
    d_bias = torch.from_numpy(inp["bias"]).to(dev)
    d_w = torch.from_numpy(inp["w"]).to(dev)
    d_out = torch.zeros(N * (2 + NB), dtype=torch.float64, device=dev)    # [ll | g_bias | g_w]
    d_ll, d_gb, d_gw = d_out[:N], d_out[N:2 * N], d_out[2 * N:]

    def step_kernel():
        ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, nlin, 0, N, path,
                       d_ll.data_ptr(), d_gb.data_ptr(), d_gw.data_ptr(), stream.cuda_stream)

    def step_dev():
        if comm is not None:                # evaluation + sum of the time shards' partial ll / gradients in one call: the
            # fused kernel's final reduction is the collective (csrc/llgrad_tc.cu, tc_final2_allreduce_kernel)
            ds.ll_grad_allreduce_dev(comm, d_bias.data_ptr(), d_w.data_ptr(), 0, 0, nlin, path, d_out.data_ptr(), stream.cuda_stream)
            return
        step_kernel()
        if world > 1:
            dist.all_reduce(d_out)

    step_kernel()                            # first call builds the split planes (part of ingest)
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t_ing0

    # host-buffer path (e2e): pinned parameter upload, eval, result download every step
    h_bias = torch.from_numpy(inp["bias"]).pin_memory()
    h_w = torch.from_numpy(inp["w"]).pin_memory()
    h_out = torch.empty(N * (2 + NB), dtype=torch.float64).pin_memory()
    e_bias = torch.empty_like(d_bias)
    e_w = torch.empty_like(d_w)
    e_out = torch.zeros_like(d_out)

    def step_e2e_host():
        """N = 1: the host-buffer C-ABI call on page-locked caller buffers (parameters in, ll / gradients out)."""
        ds.ll_grad_host_ptrs(h_bias.data_ptr(), h_w.data_ptr(), 0, 0, nlin, 0, N, path,
                             h_out[:N].data_ptr(), h_out[N:2 * N].data_ptr(), h_out[2 * N:].data_ptr())

    def step_e2e():
        if world == 1:
            return step_e2e_host()
        e_bias.copy_(h_bias, non_blocking=True)
        e_w.copy_(h_w, non_blocking=True)
        if comm is not None:
            ds.ll_grad_allreduce_dev(comm, e_bias.data_ptr(), e_w.data_ptr(), 0, 0, nlin, path, e_out.data_ptr(), stream.cuda_stream)
        else:
            ds.ll_grad_dev(e_bias.data_ptr(), e_w.data_ptr(), 0, 0, nlin, 0, N, path,
                           e_out[:N].data_ptr(), e_out[N:2 * N].data_ptr(), e_out[2 * N:].data_ptr(),
                           stream.cuda_stream)
            if world > 1:
                dist.all_reduce(e_out)
        h_out.copy_(e_out, non_blocking=True)
        stream.synchronize()                # the caller reads ll / gradient on the host every step

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, dev_blocks = timer.blocks(step_dev, args.steps, min_total_s)
    # dominant-kernel timing for the roofline: the ll+grad launch sequence alone (no collective), same clock window
    ms_kernel = (timer.blocks(step_kernel, args.steps, min_total_s)[0] if world > 1 else ms_dev) / args.steps
    for _ in range(3):
        step_e2e()
    if world == 1:                          # the call synchronises on the handle's own stream: wall clock is the measure
        ms_e2e, e2e_blocks = wall_blocks(step_e2e, args.steps, torch.cuda.synchronize, min_total_s)
        if not np.array_equal(h_out[:N].numpy(), d_ll.cpu().numpy()):
            raise SystemExit("host-buffer call and device-resident call disagree")
    else:
        ms_e2e, e2e_blocks = timer.blocks(step_e2e, args.steps, min_total_s)
    ms_coll = None
    if world > 1:                           # the collective alone, same buffer: latency at 29 KB, bandwidth at 671 MB
        def coll():
            if comm is not None:
                comm.allreduce_sum_dev(d_out.data_ptr(), d_out.data_ptr(), d_out.numel(), stream.cuda_stream)
            else:
                dist.all_reduce(d_out)
        ms_coll = timer.blocks(coll, 5, 0.0)[0] / 5
    clocks = sampler.stop() if rank == 0 else None

    # the same evaluation through the host-buffer C-ABI call a Python user makes (numpy in, numpy out; the library
    # stages through its own pinned buffers and replays a CUDA graph): wall clock, single GPU only
    host_entry = None
    if world == 1:
        for _ in range(3):
            ds.ll_grad(inp["bias"], inp["w"], nlin=nlin, path=path)
        ms_he, _ = wall_blocks(lambda: ds.ll_grad(inp["bias"], inp["w"], nlin=nlin, path=path), args.steps,
                               torch.cuda.synchronize, min(min_total_s, 0.3))
        host_entry = args.steps / (ms_he * 1e-3)

    ll_host = d_ll.cpu().numpy()
    if not np.all(np.isfinite(ll_host)):
        raise SystemExit("non-finite log-likelihood in the benchmark run")

    line = None
    if rank == 0:
        roof, info = llgrad_roofline(ds, wl, ms_kernel, path, workload_key)
        per_step = ms_dev / args.steps
        h2d = (N + N * NB) * 8
        d2h = N * (2 + NB) * 8
        line = {
            "metric": METRIC, "value": world / (per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": info.get("dtype", "f64"), "data": "synthetic",
            "config": workload_config(wl),
            "details": {"path": info.get("name", path) + (" (from the spikes: K1 per time chunk inside every evaluation)" if wl.get("x_dtype") == "none" else ""),
                        "x_bytes_per_gpu": T * ds.ldx * 4, "x_resident": wl.get("x_dtype") != "none",
                        "sharding": "time-sharded: one T-bin sequence per GPU, allreduce(sum) of ll/grad partials"
                        if world > 1 else "single GPU",
                        "collective": collective, "ingest_s_incl_filter_and_planes": ingest_s,
                        "timing": "median of %d blocks of %d steps (each block: barrier + synchronize both sides, CUDA events, "
                                  "max over ranks); timed total %.2f s" % (len(dev_blocks), args.steps, sum(dev_blocks) * 1e-3),
                        "block_ms_min_max": [min(dev_blocks), max(dev_blocks)]},
            "collective_alone": None if ms_coll is None else {
                "ms": ms_coll, "bytes": int(d_out.numel()) * 8,
                "bus_GBps": 2.0 * (world - 1) / world * d_out.numel() * 8 / (ms_coll * 1e-3) / 1e9,
                "reference": "8-rank NCCL all-reduce bus bandwidth 725 GB/s at 1 GiB (B200_PROFILING.md); SURVEY 5 ring estimate 1.3 ms for 671 MB"},
            "clocks": clocks,
            "e2e": {"value": world / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": ("pyglm_b200_ll_grad on page-locked host buffers: parameter upload, evaluation and download of "
                             "ll / gradients replayed as one CUDA graph, one synchronise per step") if world == 1 else
                            ("pyglm_b200_ll_grad_allreduce_dev (evaluation + sum over ranks) between a pinned-host upload of the parameters and a "
                             "pinned-host download of ll / gradients, one stream synchronise per step"),
                    "blocks": len(e2e_blocks), "host_entry_value": host_entry,
                    "host_entry_call": "pyglm_b200_ll_grad with pageable numpy arrays in and out (staged by the library), wall clock"},
            "gpu_launches": int(info.get("launches_per_eval", 5)) * args.steps,
            "roofline": roof,
        }
        if world == 1 and cpu_budget_s > 0:              # CPU baseline on a bounded sample (rank 0, N=1 only)
            line["cpu_baseline"] = llgrad_cpu_baseline(inp, wl, cpu_budget_s)
    ds.close()
    del d_bias, d_w, d_out, e_bias, e_w, e_out
    torch.cuda.empty_cache()
    return line


def fromspikes_record(args, wl, pg, torch, dist, timer, world, rank, local_rank, shard_of=1):
    """ll + gradient from the spike trains (x_dtype="none"): every evaluation runs K1 into chunk-sized operand planes and
    the tensor-core kernels on them; nothing but the spikes is resident.  The rank evaluates the postsynaptic neurons
    neuron_shard(N, shard_of or world, rank) -- with several GPUs that is north_star's split of C4 by postsynaptic neuron:
    the same recording on every GPU, no collective inside an evaluation (strong scaling over neurons)."""
    from theano_pyglm_b200.utils.parallel_util import neuron_shard
    N, T, B = wl["N"], wl["T"], wl["B"]
    NB = N * B
    dev, stream = timer.dev, timer.stream
    parts = world if world > 1 else shard_of
    n_lo, n_hi = neuron_shard(N, parts, rank if world > 1 else 0)
    nc = n_hi - n_lo
    inp = make_inputs(wl, seed=1234)                          # the same recording on every rank
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], device=local_rank, x_dtype="none")
    d_bias = torch.from_numpy(inp["bias"]).to(dev)
    d_w = torch.from_numpy(inp["w"]).to(dev)
    d_out = torch.zeros(nc * (2 + NB), dtype=torch.float64, device=dev)
    step = lambda: ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, "explinear", n_lo, n_hi, "auto", d_out[:nc].data_ptr(),
                                  d_out[nc:2 * nc].data_ptr(), d_out[2 * nc:].data_ptr(), stream.cuda_stream)
    for _ in range(2):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    steps = max(2, min(args.steps, 5))
    ms, blocks = timer.blocks(step, steps, 0.3)
    ms /= steps
    clocks = sampler.stop() if rank == 0 else None
    # the host-buffer call for the same neurons (parameters up, ll / gradients down): e2e
    ms_e2e = None
    if world == 1:
        call = lambda: ds.ll_grad(inp["bias"], inp["w"], n_lo=n_lo, n_hi=n_hi)
        for _ in range(2):
            call()
        ms_e2e = wall_blocks(call, 2, torch.cuda.synchronize, 0.2)[0] / 2
    if not bool(torch.isfinite(d_out[:nc]).all()):
        raise SystemExit("non-finite log-likelihood in the from-spikes run")
    rd, wr = ds.filter_bytes()
    ds.close()
    if rank != 0:
        return None
    gemm = NB > 160
    tpeak, tsrc = load_tensor_peak()
    hpeak, hsrc = load_peaks()
    alg_flops = 4.0 * T * nc * N * B
    rec = {"metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f16x2-split/f32-acc/f64-sum", "data": "synthetic", "config": workload_config(wl), "clocks": clocks,
           "details": {"path": "from-spikes: K1 -> chunk planes -> " + ("tcgen05 GEMM kernels" if gemm else "tcgen05 fused kernel"),
                       "neurons_per_gpu": nc, "resident_bytes_per_gpu": int(inp["S"].nbytes),
                       "x_bytes_never_materialised": T * NB * 4,
                       "sharding": ("by postsynaptic neuron: %d of %d columns per GPU, the whole recording on every GPU, no collective "
                                    "inside an evaluation" % (nc, N)) if parts > 1 else "single GPU, all neurons",
                       "producer_bytes_per_eval": rd + wr, "blocks": len(blocks)},
           "gpu_launches": steps * 8}
    if ms_e2e is not None:
        rec["e2e"] = {"value": 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": (N + N * NB) * 8, "d2h_bytes_per_step": nc * (2 + NB) * 8,
                      "call": "pyglm_b200_ll_grad (numpy in / out) on the spikes-only dataset"}
    if gemm:
        rec["roofline"] = {"bound": "tensor", "achieved": alg_flops / (ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                           "frac": alg_flops / (ms * 1e-3) / 1e12 / tpeak, "traffic": None,
                           "kernel": "filter_kernel+tc_gemm_fwd_kernel+tc_gemm_bwd_kernel", "peak_source": tsrc, "kernel_ms": ms,
                           "algorithmic_flops": alg_flops, "executed_over_algorithmic": 3.0,
                           "note": "whole evaluation including the operand producer (K1), which writes and the GEMMs re-read "
                                   "%.1f GB of planes per evaluation through the chunk buffers" % (wr / 1e9)}
    else:
        alg = rd + 16 * N * NB                                 # the spikes once: the only algorithmic HBM traffic left
        rec["roofline"] = {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hpeak, "unit": "GB/s",
                           "frac": alg / (ms * 1e-3) / 1e9 / hpeak, "traffic": None, "kernel": "filter_kernel+tc_fused_kernel",
                           "peak_source": hsrc, "kernel_ms": ms, "algorithmic_bytes": alg,
                           "note": "bound by the producer's gather (shared-memory operand fetch of the FP64 sums), not by HBM: "
                                   "DESIGN.md section 4, K1"}
    return rec


def run_scale_extras(args, pg, torch, dist, timer, world, rank, local_rank):
    """Sub-records of the multi-GPU line: the C5 population time-sharded with its 671 MB all-reduce, and the collapsed
    Gibbs sweep of C3 partitioned by postsynaptic neuron (north_star's two splits).  Every rank takes part; rank 0
    returns the records."""
    import copy
    out = []
    a5 = copy.copy(args)
    a5.steps, a5.warmup = min(args.steps, 4), 2
    jobs = [("scale_c5", lambda: llgrad_record(a5, WORKLOADS["c5-shard"], "c5-shard", pg, torch, dist, timer, world, rank,
                                               local_rank, None, "nccl all_reduce (671 MB per step)", 0.0, 0.0)),
            ("scale_c5_b10", lambda: llgrad_record(a5, WORKLOADS["c5-shard-b10"], "c5-shard-b10", pg, torch, dist, timer, world, rank,
                                                   local_rank, None, "nccl all_reduce (1.34 GB per step)", 0.0, 0.0)),
            ("scale_gibbs", lambda: gibbs_sharded_record(args, WORKLOADS["c3-gibbs"], pg, torch, dist, timer, world, rank, local_rank)),
            ("scale_c4_neuron", lambda: fromspikes_record(args, WORKLOADS["c4-neuron-shard"], pg, torch, dist, timer, world, rank, local_rank))]
    for name, job in jobs:
        t0 = time.perf_counter()
        ok = torch.ones(1, device=timer.dev)
        rec = None
        try:
            rec = job()
        except Exception as exc:
            rec = {"error": "%s: %s" % (type(exc).__name__, exc)}
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            rec = rec if rec is not None else {}
            if ok.item() < 1 and "error" not in rec:
                rec = {"error": "a peer rank failed"}
            rec["name"] = name
            rec["wall_s"] = time.perf_counter() - t0
            out.append(rec)
        torch.cuda.empty_cache()
    return out if rank == 0 else None


def gibbs_sharded_record(args, wl, pg, torch, dist, timer, world, rank, local_rank):
    """Collapsed Gibbs over A/W partitioned by postsynaptic neuron (parallel_gibbs.py:162-168): every rank holds the
    spikes of ALL presynaptic neurons (spikes-only dataset: K4 gathers its currents from the spike trains, no X), owns
    the columns neuron_shard(N, world, rank) and resamples one edge per owned column per lock-step.  No collective
    inside a sweep; the resampled columns of A (int8) and W (f64) are all-gathered over NCCL once per sweep.  Strong
    scaling: the N^2 edges of one sweep are divided among the GPUs."""
    from theano_pyglm_b200.utils.parallel_util import neuron_shard
    N, T, B = wl["N"], wl["T"], wl["B"]
    Q = 11
    dev = timer.dev
    stream = timer.stream
    inp = make_gibbs_inputs(wl, 1234)                       # the same recording and state on every rank
    n_lo, n_hi = neuron_shard(N, world, rank)
    nc = n_hi - n_lo
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], x_dtype="none", device=local_rank)
    ds.gibbs_begin(inp["bias"], inp["w"], inp["A"], inp["W"], nlin="explinear", n_lo=n_lo, n_hi=n_hi)
    rng = np.random.default_rng(7)
    orders = np.stack([rng.permutation(N) for _ in range(N)])
    xs, _ws = np.polynomial.hermite.hermgauss(10)
    cols = np.arange(n_lo, n_hi, dtype=np.int32)
    pres0 = orders[n_lo:n_hi, 0].astype(np.int32)
    diag = pres0 == cols
    cand = np.concatenate([np.sqrt(2) * np.where(diag, 0.5, 1.0)[:, None] * xs[None, :] + np.where(diag, -0.2, 0.0)[:, None],
                           np.zeros((nc, 1))], axis=1)
    d_cols, d_pres = torch.from_numpy(cols).to(dev), torch.from_numpy(pres0).to(dev)
    d_cand = torch.from_numpy(cand).to(dev)
    d_out = torch.empty((nc, Q), dtype=torch.float64, device=dev)

    def step():
        ds.gibbs_delta_ll_dev(nc, d_cols.data_ptr(), d_pres.data_ptr(), Q, d_cand.data_ptr(), d_out.data_ptr(), stream.cuda_stream)

    # the per-sweep splice: owners' columns of A and W to every rank
    wmax = -(-N // world)
    a_mine = torch.zeros((wmax, N), dtype=torch.int8, device=dev)
    w_mine = torch.zeros((wmax, N), dtype=torch.float64, device=dev)
    a_all = torch.empty((world * wmax, N), dtype=torch.int8, device=dev)
    w_all = torch.empty((world * wmax, N), dtype=torch.float64, device=dev)

    def gather():
        dist.all_gather_into_tensor(a_all, a_mine)
        dist.all_gather_into_tensor(w_all, w_mine)

    for _ in range(max(args.warmup, 3)):
        step()
    gather()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_blk, blocks = timer.blocks(step, args.steps, 0.3, max_blocks=50)
    ms_step = ms_blk / args.steps
    ms_gather = timer.blocks(gather, 5, 0.0)[0] / 5
    clocks = sampler.stop() if rank == 0 else None
    ds.gibbs_end()
    ds.close()
    if rank != 0:
        return None
    sweep_ms = N * ms_step + ms_gather
    return {"metric": "Gibbs edge-sweeps/sec", "value": 1e3 / sweep_ms, "unit": "sweeps/s", "n_gpus": world,
            "steps": args.steps, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "N": N, "T_bins_per_gpu": T, "B": B, "R": 200, "nlin": "explinear",
                       "candidates_per_edge": Q, "l2": "inputs_exceed_l2"},
            "details": {"sharding": "by postsynaptic neuron: %d columns per GPU, all presynaptic spike trains on every GPU "
                                    "(spikes-only dataset, from-spikes K4)" % wmax,
                        "step": "one lock-step = one edge per owned column x 11 candidate weights; a sweep is N steps + one splice",
                        "collective": "NCCL all_gather of A[:, n] (int8) and W[:, n] (f64) once per sweep",
                        "splice_ms": ms_gather, "sweep_ms": sweep_ms, "blocks": len(blocks),
                        "ars": "excluded (W | A=1 from the prior, as in the single-GPU record)"},
            "clocks": clocks, "gpu_launches": 2 * args.steps}


def filter_record(args, pg, torch, timer, local_rank):
    """K1 alone on the C2 recording: spikes resident; one pass rewrites what ingest produces for the default dataset --
    the FP32 filtered spike train AND the FP16 split planes of the tensor-core path (single-pass ingest)."""
    wl = WORKLOADS["c2"]
    N, T, B = wl["N"], wl["T"], wl["B"]
    inp = make_inputs(wl, 1234)
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], device=local_rank)
    stream = timer.stream
    step = lambda: ds.refilter(stream.cuda_stream)
    for _ in range(3):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, blocks = timer.blocks(step, args.steps, 0.3)
    clocks = sampler.stop()
    ms /= args.steps
    # e2e: host spikes in (pageable numpy), filtered spike train resident on return: dataset_create
    e2e_ms = []
    for _ in range(7):
        t0 = time.perf_counter()
        pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], device=local_rank).close()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    ms_e2e = float(np.median(e2e_ms))                 # median of 7 (device allocations make the first ones slow)
    peak, peak_src = load_peaks()
    rd, wr = ds.filter_bytes()
    alg = rd + wr
    rec = {"name": "k1-filter-c2", "metric": "spike-history filter passes/sec", "value": 1e3 / ms, "unit": "passes/s",
           "ms_per_step": ms, "steps": args.steps, "blocks": len(blocks), "dtype": "u8 in, f64 accumulate, f32 + f16x2-split planes out",
           "config": dict(workload_config(wl), workload="K1 on " + wl["desc"]), "clocks": clocks,
           "e2e": {"value": 1e3 / ms_e2e, "unit": "passes/s", "h2d_bytes_per_step": int(inp["S"].nbytes), "d2h_bytes_per_step": 0,
                   "call": "pyglm_b200_dataset_create: upload of the uint8 spikes, K1, spike transpose; synchronous"},
           "gpu_launches": args.steps,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": _traffic("filter_kernel", "c2"), "kernel": "filter_kernel",
                        "peak_source": peak_src, "kernel_ms": ms, "algorithmic_bytes": alg}}
    if not args.no_cpu:
        from oracle import pyglm_oracle as orc
        T_s = 200_000
        Ssub = inp["S"][:T_s].astype(np.float64)
        t0 = time.perf_counter()
        orc.convolve_with_basis(Ssub, inp["ibasis"])          # the reference's own call (B FFT convolutions)
        t_cpu = (time.perf_counter() - t0) * T / T_s
        rec["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "passes/s", "cores": 1, "kind": "port",
                               "sample": "scipy.signal.fftconvolve form of utils/basis.py:201-236 on the first %d of %d bins, scaled" % (T_s, T)}
    ds.close()
    return rec


def run_extras(args, pg, torch, dist, timer, local_rank):
    """The rest of BASELINE.json's metric in the same line: K1, C3 ll+grad, C3 Gibbs sweeps/s, a C4 time shard."""
    out = []
    jobs = [("k1-filter-c2", lambda: filter_record(args, pg, torch, timer, local_rank)),
            ("c3", lambda: llgrad_record(args, WORKLOADS["c3"], "c3", pg, torch, dist, timer, 1, 0, local_rank, None,
                                         "none (single GPU)", 0.3, 0.0 if args.no_cpu else 6.0)),
            ("c3-gibbs", lambda: gibbs_record(args, WORKLOADS["c3-gibbs"], "c3-gibbs", 0.3)),
            ("c4-shard", lambda: llgrad_record(args, WORKLOADS["c4-shard"], "c4-shard", pg, torch, dist, timer, 1, 0, local_rank,
                                               None, "none (single GPU)", 0.3, 0.0 if args.no_cpu else 6.0)),
            ("c4-shard-from-spikes", lambda: llgrad_record(args, WORKLOADS["c4-shard-from-spikes"], "c4-shard-from-spikes", pg, torch, dist,
                                                           timer, 1, 0, local_rank, None, "none (single GPU)", 0.3, 0.0)),
            ("c2-from-spikes", lambda: fromspikes_record(args, WORKLOADS["c2-from-spikes"], pg, torch, dist, timer, 1, 0, local_rank)),
            ("c4-neuron-shard", lambda: fromspikes_record(args, WORKLOADS["c4-neuron-shard"], pg, torch, dist, timer, 1, 0, local_rank,
                                                          shard_of=8))]
    for name, job in jobs:
        t0 = time.perf_counter()
        try:
            rec = job()
            rec["name"] = name
            rec["wall_s"] = time.perf_counter() - t0
        except Exception as exc:                          # an extra must never cost the headline line
            rec = {"name": name, "error": "%s: %s" % (type(exc).__name__, exc)}
        out.append(rec)
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------
# Gibbs workload: edge-sweeps/sec (one sweep = N^2 edges x 11 delta-ll candidates, ARS excluded)
# --------------------------------------------------------------------------------------
def make_gibbs_inputs(wl, seed):
    """Dirichlet impulses (beta rows sum to 1), normalised basis, ER graph with the
    stabilised sparsity, Gaussian weights with a refractory diagonal, bias ~ N(20, 0.25^2), softplus."""
    from theano_pyglm_b200.utils.basis import create_basis, interpolate_basis
    N, T, B = wl["N"], wl["T"], wl["B"]
    rng = np.random.default_rng(seed)
    S = (rng.random((T, N)) < 0.02).astype(np.uint8)
    prms = dict(type='cosine', n_eye=0, n_cos=B, a=1.0 / 120, b=0.5, orth=False, norm=True)
    ib = interpolate_basis(create_basis(prms), 0.001, 0.2, True, "dirichlet")
    g = rng.gamma(1.0, 1.0, size=(N, N, B))
    w = (g / g.sum(axis=2, keepdims=True)).reshape(N, N * B)
    rho = min(1.0, (0.7 + 0.2) ** 2 / N)
    A = (rng.random((N, N)) < rho).astype(np.int8)
    np.fill_diagonal(A, 1)
    W = rng.standard_normal((N, N))
    W[np.diag_indices(N)] = -0.2 + 0.5 * rng.standard_normal(N)
    bias = 20.0 + 0.25 * rng.standard_normal(N)
    return dict(S=S, ibasis=ib, bias=bias, w=w, A=A, W=W, dt=0.001, rho=rho)


def gibbs_record(args, wl, workload_key, min_total_s=0.5):
    """Collapsed Gibbs over A/W at one GPU: the batched delta-ll kernel (device-timed `value`), and one lock-step of the
    sweep through the host-buffer calls (`e2e`: candidates in, 11 log-likelihoods out, host decision rule, commit).
    ARS probes for W | A = 1 are NOT part of either figure (each is one more Q = 1 delta-ll launch; hips' ARS is
    un-vendored, so their number per accepted edge is not defined by the reference tree): W comes from the prior."""
    import torch
    import theano_pyglm_b200 as pg
    from scipy.special import logsumexp
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    N, T, B = wl["N"], wl["T"], wl["B"]
    Q = 11
    inp = make_gibbs_inputs(wl, 1234)
    dev = torch.device("cuda", 0)
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], x_dtype="f64", device=0)
    ds.gibbs_begin(inp["bias"], inp["w"], inp["A"], inp["W"], nlin="explinear")
    rng = np.random.default_rng(7)
    orders = np.stack([rng.permutation(N) for _ in range(N)])          # shuffled presynaptic order per column
    xs, ws = np.polynomial.hermite.hermgauss(10)
    log_wgh = np.log(ws / np.sqrt(np.pi))
    cols = np.arange(N, dtype=np.int32)

    def candidates(pres):
        diag = pres == cols
        mu = np.where(diag, -0.2, 0.0)[:, None]
        sig = np.where(diag, 0.5, 1.0)[:, None]
        return np.concatenate([np.sqrt(2) * sig * xs[None, :] + mu, np.zeros((N, 1))], axis=1), mu[:, 0], sig[:, 0]

    pA = np.where(np.eye(N, dtype=bool), 1.0 - 1e-8, inp["rho"])
    counter = [0]

    def step_e2e():
        """One lock-step of the sweep: every column resamples its s-th edge (host buffers in and out)."""
        s = counter[0]
        counter[0] += 1
        pres = orders[:, s % N].astype(np.int32)
        cand, mu, sig = candidates(pres)
        ll = ds.gibbs_delta_ll(cols, pres, cand)                        # (N, 11)
        log_G = logsumexp(ll[:, :10] + log_wgh[None, :], axis=1)       # gibbs.py:1015-1022
        lp_A = np.log(pA[pres, cols]) + log_G
        lp_no = np.log1p(-pA[pres, cols]) + ll[:, 10]
        p_no = np.exp(lp_no - np.logaddexp(lp_no, lp_A))
        a_new = (rng.random(N) > p_no).astype(np.int8)                  # log_sum_exp.py:26-32
        w_new = mu + sig * rng.standard_normal(N)                       # prior draw; ARS is not part of the metric
        ds.gibbs_commit(cols, pres, a_new, w_new)

    # device-timed: the batched delta-ll kernel alone, operands resident
    d_cols = torch.from_numpy(cols).to(dev)
    pres0 = orders[:, 0].astype(np.int32)
    d_pres = torch.from_numpy(pres0).to(dev)
    d_cand = torch.from_numpy(candidates(pres0)[0]).to(dev)
    d_out = torch.empty((N, Q), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    timer = Timer(torch, None, 1, dev, stream)

    def step_dev():
        ds.gibbs_delta_ll_dev(N, d_cols.data_ptr(), d_pres.data_ptr(), Q, d_cand.data_ptr(), d_out.data_ptr(),
                              stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ms_blk, blocks = timer.blocks(step_dev, args.steps, min_total_s, max_blocks=50)
    ms_dev = ms_blk / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e_blk, _ = wall_blocks(step_e2e, args.steps, torch.cuda.synchronize, min_total_s, max_blocks=20)
    ms_e2e = ms_e2e_blk / args.steps
    clocks = sampler.stop()

    peak, peak_src = load_peaks()
    alg_bytes = N * T * (8 + B * 8 + 1)          # per edge and bin: I_net f64 + X slice f64 + spike byte
    achieved = alg_bytes / (ms_dev * 1e-3) / 1e9
    gibbs_traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            gibbs_traffic = json.load(f).get("gibbs_delta_kernel", {}).get(workload_key)
    # FP64 work per bin: Q softplus evaluations of 21 fused operations + the u / base current (B + 2)
    fp64_flops = 2.0 * N * T * (Q * 21 + B + 2)
    try:
        fp64_peak, fp64_src = pg.engine.measure_fp64_peak(0), "measured here: register-resident DFMA loop (pyglm_b200_measure_fp64_peak)"
    except Exception:
        fp64_peak, fp64_src = 40.0, "nominal B200 FP64 (probe unavailable)"
    line = {
        "metric": "Gibbs edge-sweeps/sec", "value": 1.0 / (N * ms_dev * 1e-3), "unit": "sweeps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "N": N, "T_bins_per_gpu": T, "B": B, "R": 200, "nlin": "explinear",
                   "candidates_per_edge": Q, "l2": "inputs_exceed_l2"},
        "details": {"step": "one lock-step = N edges (one per column) x 11 candidate weights; a sweep is N steps",
                    "edges_per_s": N / (ms_dev * 1e-3), "blocks": len(blocks),
                    "ars": "excluded: W | A=1 is drawn from the prior in this benchmark (hips ARS is un-vendored; each probe "
                           "would be one more Q=1 launch of the same kernel)"},
        "clocks": clocks,
        "e2e": {"value": 1.0 / (N * ms_e2e * 1e-3), "unit": "sweeps/s",
                "h2d_bytes_per_step": N * (4 + 4 + Q * 8) + N * (4 + 4 + 1 + 8), "d2h_bytes_per_step": N * Q * 8,
                "includes": "host decision rule (logsumexp, Bernoulli draw) and gibbs_commit"},
        "gpu_launches": 2 * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": gibbs_traffic, "kernel": "gibbs_delta_kernel", "peak_source": peak_src, "kernel_ms": ms_dev,
                     "algorithmic_bytes": alg_bytes,
                     "note": "FP64-ALU bound, not HBM (SURVEY.md 8d): the second figure is the FP64 rate",
                     "fp64": {"achieved": fp64_flops / (ms_dev * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                              "frac": fp64_flops / (ms_dev * 1e-3) / 1e12 / fp64_peak, "peak_source": fp64_src}},
    }
    if not args.no_cpu:
        from oracle import pyglm_oracle as orc
        T_s = min(T, 100_000)
        Ssub = inp["S"][:T_s].astype(np.float64)
        fS_pre = orc.convolve_with_basis_direct(Ssub[:, :1], inp["ibasis"])          # one presynaptic column is enough
        u = fS_pre[:, 0, :] @ inp["w"][0].reshape(N, B)[0]
        I_other = 0.3 * u
        cand = candidates(pres0)[0][0]
        t0 = time.perf_counter()
        for wq in cand:
            orc.gibbs_glm_ll(inp["bias"][0], 0.0, I_other, u, wq, Ssub[:, 0], inp["dt"], orc.NLIN_SOFTPLUS)
        t_edge = (time.perf_counter() - t0) * (T / T_s)
        line["cpu_baseline"] = {"value": 1.0 / (t_edge * N * N), "unit": "sweeps/s", "cores": 1, "kind": "port",
                                "sample": "11 delta-ll evaluations (gibbs.py:910-937 restated, numpy float64) for one "
                                          "edge on the first %d of %d bins, scaled to N^2 edges" % (T_s, T)}
    ds.gibbs_end()
    ds.close()
    torch.cuda.empty_cache()
    return line


def run_gibbs(args, wl):
    print(json.dumps(gibbs_record(args, wl, args.workload)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "fp64", "tc"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads of the default (c2, one GPU) line")
    ap.add_argument("--allreduce", default="p2p", choices=["p2p", "nccl"],
                    help="collective of the time-sharded run: peer-memory one-shot kernel (default) or NCCL")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl.get("gibbs"):
        run_gibbs(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
