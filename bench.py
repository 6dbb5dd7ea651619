#!/usr/bin/env python
"""Benchmark of the OverIVA demixing loop on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Metric (BASELINE.json): mixture-seconds separated per second at 20 iterations.  One *step* is one complete
``overiva`` call (relayout + input covariance + init + 20 epochs + final demix with projection back) over
one batch of synthetic mixtures.  Workload at every N: BASELINE config 4's per-GPU share -- 512
independent mixtures of 15 s at 16 kHz (T=116, F=2049, M=6 microphones, K=2 sources, laplace model,
complex128), one shard per rank, no collective in the data path (weak scaling).

Printed JSON (one line, rank 0): ``value`` = device-resident throughput (X already in HBM), ``e2e`` = the
same metric through the public ``overiva_batch`` call with pinned HOST buffers (H2D of X and D2H of Y inside
the timed region), ``e2e_audio`` = the same workload entered as time-domain AUDIO through
``overiva_b200.stft.separate_batch`` (STFT / iSTFT on the device: half the PCIe bytes), ``roofline`` for the dominant kernel (weighted covariance) from CUDA events recorded
around its launches inside the timed region, ``cpu_baseline`` = the numpy oracle (a port of the reference's
algorithm) on this box's host cores over a bounded sample of the same workload.

``--impl reference`` times the reference's CPU algorithm (the oracle port; the reference itself is pure
Python/numpy and does not travel to the GPU box) with all host cores -- a pool of one-BLAS-thread workers,
the reference's own ipyparallel + mkl.set_num_threads(1) design (overiva_sim.py:85-91).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json config 4 (one GPU's share): 512 x (T=116, F=2049, M=6), K=2, laplace, 20 iterations
FS = 16000
DURATION_S = 15.0
T, F, M, K = 116, 2049, 6, 2
N_ITER = 20
MODEL = "laplace"
BATCH_PER_GPU = 512
METRIC = "mixture-seconds separated per second (20 iterations)"
UNIT = "mixture-s/s"


def workload_config(batch, n_gpus):
    return {
        "workload": "BASELINE cfg4 shard: %d independent mixtures/GPU, 15 s @ 16 kHz, STFT 4096/2048 -> "
        "(T=%d,F=%d,M=%d), overiva K=%d %s, n_iter=%d, proj_back, complex128" % (batch, T, F, M, K, MODEL, N_ITER),
        "mixtures_per_gpu": batch,
        "global_mixtures": batch * n_gpus,
        "n_iter": N_ITER,
        "parallelism": "batch split across %d GPU(s), no data-path collective" % n_gpus,
        "l2": "inputs (%.1f GB per GPU) are far larger than the 126 MB L2; no flush needed" % (batch * T * F * M * 16 / 1e9),
    }


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's algorithm)
# --------------------------------------------------------------------------------------------------
_CPU_X = {}


def _cpu_worker(args):
    seed, shape, n_iter = args
    import numpy as np
    from threadpoolctl import threadpool_limits

    from oracle import overiva_oracle as orc
    from overiva_b200.synth import stft_domain_mixture

    if "X" not in _CPU_X:  # each worker draws ONE mixture (at warm-up) and re-separates it: input generation
        _CPU_X["X"] = stft_domain_mixture(seed, shape[0], shape[1], shape[2], K)  # is not part of the timed call
    X = _CPU_X["X"]
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        Y = orc.overiva(X, n_src=K, n_iter=n_iter, proj_back=True, model=MODEL)
        dt = time.perf_counter() - t0
    assert np.all(np.isfinite(Y))
    return dt


class CpuPool:
    """`cores` worker processes, one BLAS thread each: the reference's own way of using a multi-core host
    (ipyparallel engines + mkl.set_num_threads(1), overiva_sim.py:85-91, rrtools/dumbparallel.py:253-279)."""

    def __init__(self, cores):
        import multiprocessing as mp

        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores)
        # warm the workers (imports, input generation for seed w, first touch) outside any timing
        self.pool.map(_cpu_worker, [(10_000 + i, (T, F, M), 1) for i in range(cores)], chunksize=1)

    def throughput(self, n_mixtures):
        jobs = [(10_000 + i % self.cores, (T, F, M), N_ITER) for i in range(n_mixtures)]
        t0 = time.perf_counter()
        per = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return n_mixtures * DURATION_S / wall, wall, per

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_step = cores  # one mixture per worker per step (a few seconds of wall time per step)
    pool = CpuPool(cores)
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        v, wall, _ = pool.throughput(per_step)
        if i >= args.warmup:
            vals.append(v)
            walls.append(wall)
    pool.close()
    value = len(vals) * per_step * DURATION_S / sum(walls)
    sample = "%d mixtures of the workload shape per step (one per worker), %d workers x 1 BLAS thread" % (per_step, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(BATCH_PER_GPU if args.batch is None else args.batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import overiva_b200 as ob
    from overiva_b200 import _lib as L
    from overiva_b200.core import DemixPlan
    from overiva_b200.synth import stft_domain_batch_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = BATCH_PER_GPU if args.batch is None else args.batch
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------
    X = stft_domain_batch_torch(B, T, F, M, K, seed=1234 + rank, device=dev)
    plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    Y = torch.empty((B, T, F, K), dtype=torch.complex128, device=dev)

    def step():
        plan.load(X)
        plan.init(L.INIT_EYE)
        plan.iterate(N_ITER)
        plan.output(True, out=Y)

    for _ in range(warmup):
        step()
    plan.raise_on_failure()
    barrier()
    l0 = plan.launches
    plan.enable_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    w1 = time.time()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = plan.launches - l0
    timing = plan.read_timing()
    plan.enable_timing(False)
    plan.raise_on_failure()
    assert bool(torch.isfinite(Y.real).all())
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    ms_per_step = ms_max / steps
    value = world * B * DURATION_S / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (weighted covariance), live CUDA-event timings ----------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cov_ms, cov_n = timing["cov"]
    alg_bytes_cov = B * F * T * M * 16  # one pass over X per launch (F*T*M*c per pass)
    achieved = alg_bytes_cov / (cov_ms / cov_n * 1e-3) / 1e9 if cov_n else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "cov_traffic.json"))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {
        "kernel": "k_cov (weighted covariance, all K sources, one pass over X)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_cov, "launches_timed": cov_n,
        "avg_launch_ms": cov_ms / cov_n if cov_n else None,
        "kernel_ms_per_step": {k: v[0] / steps for k, v in timing.items()},
        "step_algorithmic_bytes": ((2 * N_ITER + 2) * F * T * M * 16 + F * T * K * 16) * B,
        "step_achieved_gbs": ((2 * N_ITER + 2) * F * T * M * 16 + F * T * K * 16) * B / (ms_per_step * 1e-3) / 1e9,
    }

    # ---- end-to-end arm: public API, pinned host buffers in, host buffers out ---------------------
    del plan
    torch.cuda.empty_cache()
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        Xh = torch.empty((B, T, F, M), dtype=torch.complex128, pin_memory=True)
        Xh.copy_(X)
        del X, Y
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(steps, 3))
        Yh = torch.empty((B, T, F, K), dtype=torch.complex128, pin_memory=True)  # caller-owned result buffer
        for _ in range(2):
            ob.overiva_batch(Xh, n_src=K, n_iter=N_ITER, proj_back=True, model=MODEL, out=Yh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ob.overiva_batch(Xh, n_src=K, n_iter=N_ITER, proj_back=True, model=MODEL, out=Yh)
        barrier()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * B * DURATION_S / (float(t_e.item()) / e2e_steps), "unit": UNIT,
            "h2d_bytes_per_step": Xh.numel() * 16, "d2h_bytes_per_step": Yh.numel() * 16,
            "steps": e2e_steps,
            "api": "overiva_b200.overiva_batch(pinned CPU tensor, out=pinned CPU tensor): chunks of 32 mixtures, "
                   "H2D / loop / D2H overlapped on three streams",
        }
        assert bool(torch.isfinite(Yh.real).all())
    except RuntimeError as exc:  # e.g. not enough pinnable host memory
        e2e = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    # ---- end-to-end from AUDIO: STFT and iSTFT on the device, only audio crosses PCIe ---
    e2e_audio = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        from overiva_b200 import stft as gstft
        from overiva_b200.synth import audio_batch_torch

        del Xh, Yh
        torch.cuda.empty_cache()
        if hasattr(torch._C, "_host_emptyCache"):  # give the 15.6 GB of pinned spectra buffers back to the OS
            torch._C._host_emptyCache()
        n_samples = int(DURATION_S * FS)
        xa = audio_batch_torch(B, n_samples, M, K, seed=99 + rank, device=dev)
        assert gstft.num_frames(n_samples, 4096, 2048) == T
        xh = torch.empty((B, n_samples, M), dtype=torch.float64, pin_memory=True)
        xh.copy_(xa)
        del xa
        torch.cuda.empty_cache()
        yh = torch.empty((B, (T - 1) * 2048 + 4096, K), dtype=torch.float64, pin_memory=True)
        for _ in range(2):
            gstft.separate_batch(xh, n_src=K, n_iter=N_ITER, framesize=4096, model=MODEL, out=yh)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            gstft.separate_batch(xh, n_src=K, n_iter=N_ITER, framesize=4096, model=MODEL, out=yh)
        barrier()
        t_a = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_a, op=dist.ReduceOp.MAX)
        e2e_audio = {
            "value": world * B * DURATION_S / (float(t_a.item()) / e2e_steps), "unit": UNIT,
            "h2d_bytes_per_step": xh.numel() * 8, "d2h_bytes_per_step": yh.numel() * 8, "steps": e2e_steps,
            "api": "overiva_b200.stft.separate_batch(pinned host audio (B,N,M) float64, out=pinned): STFT 4096/2048 "
                   "-> loop -> iSTFT on the device, chunks of 32 mixtures on three streams",
        }
        assert bool(torch.isfinite(yh).all())
        del xh, yh
    except (RuntimeError, NameError) as exc:
        e2e_audio = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n = 2 * cores
        pool = CpuPool(cores)
        v, wall, per = pool.throughput(n)
        pool.close()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d mixtures of the workload shape, %d worker processes x 1 BLAS thread, %.1f s wall, "
                         "%.2f s per mixture per core" % (n, cores, wall, statistics.median(per))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(B, world), "clocks": clocks,
            "e2e": e2e, "e2e_audio": e2e_audio, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="mixtures per GPU (default 512 = BASELINE cfg4 / 8)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the two end-to-end legs (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
