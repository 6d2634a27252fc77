#!/usr/bin/env python
"""Benchmark of the pyglm hot path on B200: population GLM ll+grad evals/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One "step" = one ll+gradient evaluation for all N neurons of one data sequence
(BASELINE.json metric).  Workload at one GPU: config C2
(standard_glm, N=27, T=10^6 bins, B=5, R=200), synthetic Bernoulli(0.02) spikes.
With --gpus G (torchrun, one rank per GPU) every rank owns its own T=10^6-bin sequence
(time-sharded, weak scaling) and the per-step partial sums are all-reduced over NCCL,
exactly as the reference sums ll over data sequences (coord_descent.py:52-57).

Prints ONE JSON line on rank 0 (contract in the task statement).
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
}
METRIC = "GLM ll+grad evals/sec"
UNIT = "evals/s"


def make_inputs(wl, seed):
    """Synthetic C2-style inputs: Bernoulli(0.02) spikes, raised-cosine basis,
    bias ~ N(20, 0.1^2), w_ir ~ N(0, 0.05^2), complete graph, unit weights, softplus."""
    from theano_pyglm_b200.utils.basis import make_standard_ibasis
    N, T, B = wl["N"], wl["T"], wl["B"]
    rng = np.random.default_rng(seed)
    S = (rng.random((T, N)) < 0.02).astype(np.uint8)
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


def cpu_eval_time(inp, wl, T_sample, repeats, shape="gemm"):
    """Best-of-`repeats` seconds for one CPU eval on the first T_sample bins."""
    fn = cpu_prepare(inp, wl, T_sample, shape)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inp = make_inputs(wl, seed=1234)
    T_sample = min(wl["T"], 100_000)
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
    sample = ("first %d of %d bins per step, whole-population BLAS GEMM form of the float64 oracle "
              "(filter excluded, like our arm); evals/s scaled by %d/%d" % (T_sample, wl["T"], T_sample, wl["T"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_step * scale * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# Our arm
# --------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, T, B = wl["N"], wl["T"], wl["B"]
    NB = N * B
    inp = make_inputs(wl, seed=1234 + rank)      # every rank owns its own sequence (time shard)
    t_ing0 = time.perf_counter()
    ds = pg.Dataset(inp["S"], inp["dt"], inp["ibasis"], device=local_rank)
    ingest_s = time.perf_counter() - t_ing0
    path = args.path
    nlin = "explinear"

    # resident parameters + outputs (device-timed path)
    d_bias = torch.from_numpy(inp["bias"]).to(dev)
    d_w = torch.from_numpy(inp["w"]).to(dev)
    d_out = torch.zeros(N * (2 + NB), dtype=torch.float64, device=dev)    # [ll | g_bias | g_w]
    d_ll, d_gb, d_gw = d_out[:N], d_out[N:2 * N], d_out[2 * N:]
    stream = torch.cuda.current_stream()

    def step_dev():
        ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, nlin, 0, N, path,
                       d_ll.data_ptr(), d_gb.data_ptr(), d_gw.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_out)          # sum of the time shards' partial ll / gradients

    # host-buffer path (e2e): pinned parameter upload, eval, result download every step
    h_bias = torch.from_numpy(inp["bias"]).pin_memory()
    h_w = torch.from_numpy(inp["w"]).pin_memory()
    h_out = torch.empty(N * (2 + NB), dtype=torch.float64).pin_memory()
    e_bias = torch.empty_like(d_bias)
    e_w = torch.empty_like(d_w)
    e_out = torch.zeros_like(d_out)

    def step_e2e():
        e_bias.copy_(h_bias, non_blocking=True)
        e_w.copy_(h_w, non_blocking=True)
        ds.ll_grad_dev(e_bias.data_ptr(), e_w.data_ptr(), 0, 0, nlin, 0, N, path,
                       e_out[:N].data_ptr(), e_out[N:2 * N].data_ptr(), e_out[2 * N:].data_ptr(),
                       stream.cuda_stream)
        if world > 1:
            dist.all_reduce(e_out)
        h_out.copy_(e_out, non_blocking=True)
        stream.synchronize()                # the caller reads ll / gradient on the host every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # dominant-kernel timing for the roofline: the ll+grad launch sequence alone (no collective)
    def step_kernel():
        ds.ll_grad_dev(d_bias.data_ptr(), d_w.data_ptr(), 0, 0, nlin, 0, N, path,
                       d_ll.data_ptr(), d_gb.data_ptr(), d_gw.data_ptr(), stream.cuda_stream)
    ms_kernel = timed(step_kernel, args.steps) / args.steps

    ll_host = d_ll.cpu().numpy()
    if not np.all(np.isfinite(ll_host)):
        raise SystemExit("non-finite log-likelihood in the benchmark run")

    if rank == 0:
        peak, peak_src = load_peaks()
        info = ds.path_info(path) if hasattr(ds, "path_info") else {}
        x_bytes = T * ds.ldx * 4
        # algorithmic bytes per eval: X read once per contraction pass + S once + params/outputs
        passes = info.get("x_passes", 2)
        alg_bytes = passes * x_bytes + T * N + 16 * N * NB
        achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
        traffic = None          # dram__bytes_read+write of the dominant kernel, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(info.get("kernel", ""), {}).get(args.workload)
        per_step = ms_dev / args.steps
        h2d = (N + N * NB) * 8
        d2h = N * (2 + NB) * 8
        line = {
            "metric": METRIC, "value": world / (per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": info.get("dtype", "f64"), "data": "synthetic",
            "config": {"workload": wl["desc"], "N": N, "T_bins_per_gpu": T, "B": B, "R": 200, "nlin": "explinear",
                       "path": info.get("name", path), "l2": "inputs_exceed_l2 (X is %d MB per GPU)" % (x_bytes >> 20),
                       "sharding": "time-sharded: one T-bin sequence per GPU, allreduce(sum) of ll/grad partials"
                       if world > 1 else "single GPU",
                       "ingest_s_incl_filter": ingest_s},
            "clocks": clocks,
            "e2e": {"value": world / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(info.get("launches_per_eval", 5)) * args.steps,
            "roofline": {"bound": info.get("bound", "hbm"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "kernel": info.get("kernel"),
                         "peak_source": peak_src, "kernel_ms": ms_kernel,
                         "algorithmic_bytes": alg_bytes},
        }
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        if world == 1 and not args.no_cpu:
            T_s = min(T, 100_000)
            t_gemm = cpu_eval_time(inp, wl, T_s, 2, "gemm") * (T / T_s)
            T_r = min(T, 20_000)
            t_ref = cpu_eval_time(inp, wl, T_r, 1, "reference") * (T / T_r)
            line["cpu_baseline"] = {
                "value": 1.0 / t_gemm, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": "float64 oracle, whole-population BLAS GEMM form on the first %d of %d bins, scaled; "
                          "reference-shaped per-neuron loop (impulse.py:58 style, %d bins, scaled): %.4g evals/s"
                          % (T_s, T, T_r, 1.0 / t_ref),
                "reference_shaped_value": 1.0 / t_ref}
        print(json.dumps(line), flush=True)
    ds.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "fp64", "tc"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
