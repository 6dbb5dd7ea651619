#!/usr/bin/env python
"""Benchmark of the OverIVA demixing loop on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Metric (BASELINE.json): mixture-seconds separated per second at 20 iterations.  One *step* is one complete
``overiva`` call (relayout + input covariance + init + 20 epochs + final demix with projection back) over
one batch of synthetic mixtures.  Workload at every N: BASELINE config 4's per-GPU share -- 512
independent mixtures of 15 s at 16 kHz (T=116, F=2049, M=6 microphones, K=2 sources, laplace model,
complex128), one shard per rank, no collective in the data path (weak scaling).

Printed JSON (one line, rank 0):
  value          device-resident throughput (X already in HBM), CUDA events, max over ranks
  e2e            the same metric through the public ``overiva_batch`` call with pinned HOST buffers (H2D of X and D2H
                 of Y inside the timed region) -- the headline against the reference arm
  e2e_c64        the same call with complex64 TRANSPORT (X and Y cross PCIe as complex64, all arithmetic fp64 on the
                 device: BASELINE's "fp32-storage / fp64-solve" mode, parity bound 1e-3) -- half the bytes
  e2e_audio      the workload entered as time-domain AUDIO (``overiva_b200.stft.separate_batch``: STFT / iSTFT on the GPU)
  roofline       dominant kernel (weighted covariance) from CUDA events recorded around its launches in the timed region
  cpu_baseline   the reference's own CPU implementation on this box's host cores over a bounded sample (N = 1 only)
  configs        BASELINE configs 1-3 (one short mixture each) device-resident on this rank's GPU, with the relative
                 error against the CPU oracle in the same run
  cfg5_freq_sharded  BASELINE config 5 (one 10-minute 48 kHz mixture, M=16, K=4) with its frequency bins sharded over
                 the N ranks and one K x T all-reduce per epoch (strong scaling), checked in-run against a single-GPU run
  host_link      pinned H2D / D2H GB/s per rank with all ranks copying at once (what bounds e2e at N > 1)

Every optional leg is RANK-SAFE: local work happens inside try/except, collectives never do; the ranks then agree
(all-reduce MIN of an ok flag) before anything is timed or reduced, and a leg that failed anywhere is reported as
``{"value": null, "error": ..., "failed_rank": r}`` by everyone.

``--impl reference`` times the reference's CPU implementation of the path -- the UNMODIFIED reference modules
(byte-compiled into oracle/_ref by ``__graft_entry__.build()``; the numpy port if absent) -- with all host cores: a pool
of one-BLAS-thread workers, the reference's own ipyparallel + mkl.set_num_threads(1) design (overiva_sim.py:85-91).
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
import traceback

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

# BASELINE.json config 5: one 10-minute 48 kHz mixture, M=16, K=4
C5_T, C5_F, C5_M, C5_K, C5_SECS = 14061, 2049, 16, 4, 600.0
C5_Q = 16  # interfering sources of the synthetic cfg5 mixture: 4 + 16 sources >= 16 channels (well-conditioned bins)


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
# CPU arm: the reference's own implementation (oracle/_ref bytecode of the unmodified modules), else the port
# --------------------------------------------------------------------------------------------------
_CPU_X = {}


def cpu_kind():
    from oracle import reference_shim

    return "reference" if reference_shim.available() else "port"


def _cpu_worker(args):
    seed, shape, n_iter = args
    import numpy as np
    from threadpoolctl import threadpool_limits

    from oracle import reference_shim
    from overiva_b200.synth import stft_domain_mixture

    if reference_shim.available():
        separate = reference_shim.ref_overiva  # the unmodified overiva.py (numpy-1.x solve rule + projection_back stub)
    else:
        from oracle import overiva_oracle

        separate = overiva_oracle.overiva
    if "X" not in _CPU_X:  # each worker draws ONE mixture (at warm-up) and re-separates it: input generation
        _CPU_X["X"] = stft_domain_mixture(seed, shape[0], shape[1], shape[2], K)  # is not part of the timed call
    X = _CPU_X["X"]
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        Y = separate(X, n_src=K, n_iter=n_iter, proj_back=True, model=MODEL)
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
    kind = cpu_kind()
    pool = CpuPool(cores)
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        v, wall, _ = pool.throughput(per_step)
        if i >= args.warmup:
            vals.append(v)
            walls.append(wall)
    pool.close()
    value = len(vals) * per_step * DURATION_S / sum(walls)
    sample = "%d mixtures of the workload shape per step (one per worker), %d workers x 1 BLAS thread; %s" % (
        per_step, cores, "the unmodified reference overiva.py (oracle/_ref bytecode)" if kind == "reference"
        else "numpy port of overiva.py (reference bytecode absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(BATCH_PER_GPU if args.batch is None else args.batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# rank-safe plumbing
# --------------------------------------------------------------------------------------------------
class Ranks:
    """The only place collectives are issued.  Nothing here runs inside a rank-local try/except."""

    def __init__(self, torch, dist, rank, world, dev):
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def agree(self, err):
        """err: None or this rank's exception.  -> (all_ok, failed_rank, message): identical on every rank."""
        if err is not None:
            print("[bench rank %d] leg failed locally: %r" % (self.rank, err), file=sys.stderr, flush=True)
        if self.world == 1:
            return err is None, (0 if err is not None else None), (repr(err)[:300] if err is not None else None)
        t = self.torch.tensor([self.world if err is None else self.rank], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        first = int(t.item())
        if first >= self.world:
            return True, None, None
        msg = repr(err)[:300] if (err is not None and first == self.rank) else "rank %d failed (see its stderr)" % first
        return False, first, msg

    def max(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min(self, x):
        return -self.max(-float(x))

    def sum(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def guarded(fn):
    """Run LOCAL work (no collectives inside!) and hand back (result, exception)."""
    try:
        return fn(), None
    except Exception as exc:  # noqa: BLE001 -- reported through Ranks.agree
        traceback.print_exc(file=sys.stderr)
        return None, exc


def failed_leg(rk, msg, extra=None):
    out = {"value": None, "unit": UNIT, "error": msg, "failed_rank": rk}
    if extra:
        out.update(extra)
    return out


def timed_wall(R, work, reps):
    """barrier | reps x work() | barrier, wall clock, max over ranks.  A local exception inside `work` is carried out of
    the region (the closing barrier is always reached) and settled with Ranks.agree by the caller."""
    R.barrier()
    t0 = time.perf_counter()
    err = None
    try:
        for _ in range(reps):
            work()
        R.torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        traceback.print_exc(file=sys.stderr)
        err = exc
    dt_local = time.perf_counter() - t0
    R.barrier()
    return dt_local, err


class HostArena:
    """ONE pinned allocation per process, reused by every end-to-end leg (peak pinned memory = the largest leg, not
    their sum; freed explicitly at the end)."""

    def __init__(self, torch):
        self.torch = torch
        self.buf = None

    def reserve(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = self.torch.empty(int(nbytes), dtype=self.torch.uint8, pin_memory=True)

    def carve(self, specs):
        """specs: [(shape, dtype)] -> views laid out back to back (256-byte aligned)."""
        torch = self.torch
        sizes = []
        for shape, dtype in specs:
            n = 1
            for s in shape:
                n *= int(s)
            sizes.append(((n * torch.empty((), dtype=dtype).element_size() + 255) // 256) * 256)
        self.reserve(sum(sizes))
        views, off = [], 0
        for (shape, dtype), sz in zip(specs, sizes):
            n = 1
            for s in shape:
                n *= int(s)
            nb = n * torch.empty((), dtype=dtype).element_size()
            views.append(self.buf[off : off + nb].view(dtype).view(*shape))
            off += sz
        return views

    def release(self):
        self.buf = None
        if hasattr(self.torch._C, "_host_emptyCache"):
            self.torch._C._host_emptyCache()


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
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
    R = Ranks(torch, dist, rank, world, dev)
    B = BATCH_PER_GPU if args.batch is None else args.batch
    steps, warmup = args.steps, max(args.warmup, 3)
    lib = L.load()

    # ---- device-resident arm (the headline leg: a failure here is fatal by design) -----------------
    X = stft_domain_batch_torch(B, T, F, M, K, seed=1234 + rank, device=dev)
    plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    Y = torch.empty((B, T, F, K), dtype=torch.complex128, device=dev)

    def step():  # one library call: relayout + C, init, the 20 epochs, projection back + final demix (oiva_plan_run)
        plan.run(X, L.INIT_EYE, None, N_ITER, True, out=Y)

    for _ in range(warmup):
        step()
    plan.raise_on_failure()
    R.barrier()
    l0 = plan.launches
    plan.enable_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R.barrier()
    w0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    R.barrier()
    w1 = time.time()
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = plan.launches - l0
    timing = plan.read_timing()
    plan.enable_timing(False)
    plan.raise_on_failure()
    assert bool(torch.isfinite(Y.real).all())
    ms_per_step = R.max(ms) / steps
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
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "cov_traffic.json")))
        traffic = tj.get("dram_bytes_per_launch")
        traffic_src = "ncu --set full capture of one launch of this kernel at this shape (profiles/cov_traffic.json: %s); " \
                      "a recorded constant, not measured in this run" % tj.get("source", "r01b")
    except (OSError, ValueError):
        pass
    step_bytes = ((2 * N_ITER + 2) * F * T * M * 16 + F * T * K * 16) * B
    pow_ms, pow_n = timing["power"]
    pow_gbs = alg_bytes_cov / (pow_ms / pow_n * 1e-3) / 1e9 if pow_n else None
    roofline = {
        "kernel": "k_cov_sweep (weighted covariance of all K sources in one pass over X, ending per bin group in the IP "
                  "sweep of that group with the covariances still in registers: covariance + solver in one kernel)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
        # the same launch in DRAM terms: the ncu-measured bytes of one launch (X + the W_hat / C traffic of the fused
        # sweep, which the algorithmic figure of SURVEY 8(d) ignores) over the live launch time
        "dram_frac": (traffic / (cov_ms / cov_n * 1e-3) / 1e9 / peak) if (traffic and cov_n) else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_cov, "launches_timed": cov_n,
        "avg_launch_ms": cov_ms / cov_n if cov_n else None,
        "second_kernel": {"kernel": "k_demix_power (demix + norm over frequency, one pass over X)",
                          "avg_launch_ms": pow_ms / pow_n if pow_n else None, "achieved": pow_gbs,
                          "frac": (pow_gbs / peak) if pow_gbs else None, "launches_timed": pow_n},
        "kernel_ms_per_step": {k: v[0] / steps for k, v in timing.items()},
        "step_algorithmic_bytes": step_bytes,
        "step_achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
        "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
    }

    # ---- fp64 rate of this GPU (DFMA vs DMMA): the denominator of the FMA-bound shapes ------------------------
    def measure_fp64():
        out = {}
        for name, kind in (("dfma_tflops", 0), ("dmma_tflops", 1)):
            v = L.C.c_double()
            L.check(lib.oiva_fp64_peak(kind, 4096, 3, L.C.byref(v), None), "oiva_fp64_peak")
            out[name] = v.value
        return out

    fp64, err = guarded(measure_fp64)
    ok, rk, msg = R.agree(err)
    if not ok:
        fp64 = {"dfma_tflops": None, "dmma_tflops": None, "error": msg}

    # ---- end-to-end legs: public API, pinned host buffers in, host buffers out -------------------------------
    del plan
    torch.cuda.empty_cache()
    arena = HostArena(torch)
    e2e_steps = max(1, min(steps, 3))
    skip = RuntimeError("skipped (--no-e2e)") if args.no_e2e else None

    def e2e_spectra(cdtype, label):
        """One leg through overiva_batch with host tensors of dtype `cdtype`."""
        esz = 16 if cdtype == torch.complex128 else 8
        state = {}

        def prepare():
            if skip:
                raise skip
            need = B * T * F * (M + K) * esz
            avail = mem_available_bytes()
            if avail is not None and need * 1.15 > avail:
                raise MemoryError("%.1f GB of pinned host memory needed, %.1f GB available" % (need / 1e9, avail / 1e9))
            Xh, Yh = arena.carve([((B, T, F, M), cdtype), ((B, T, F, K), cdtype)])
            for b0 in range(0, B, 64):  # fill the host buffer from the device-resident workload (outside the timing)
                Xh[b0 : b0 + 64].copy_(X[b0 : b0 + 64].to(cdtype))
            torch.cuda.synchronize()
            state["Xh"], state["Yh"] = Xh, Yh
            for _ in range(2):
                ob.overiva_batch(Xh, n_src=K, n_iter=N_ITER, proj_back=True, model=MODEL, out=Yh)

        _, err = guarded(prepare)
        ok, rk, msg = R.agree(err)
        if not ok:
            return failed_leg(rk, msg)
        Xh, Yh = state["Xh"], state["Yh"]
        dt, err = timed_wall(R, lambda: ob.overiva_batch(Xh, n_src=K, n_iter=N_ITER, proj_back=True, model=MODEL, out=Yh),
                             e2e_steps)
        if err is None and not bool(torch.isfinite(torch.view_as_real(Yh)).all()):
            err = FloatingPointError("non-finite output")
        ok, rk, msg = R.agree(err)
        if not ok:
            return failed_leg(rk, msg)
        dt = R.max(dt)
        return {
            "value": world * B * DURATION_S / (dt / e2e_steps), "unit": UNIT,
            "h2d_bytes_per_step": Xh.numel() * esz, "d2h_bytes_per_step": Yh.numel() * esz, "steps": e2e_steps,
            "ms_per_step": 1e3 * dt / e2e_steps,
            "api": "overiva_b200.overiva_batch(pinned CPU %s tensor, out=pinned CPU tensor): chunks of 32 mixtures, "
                   "H2D / loop / D2H overlapped on three streams%s" % (label, "" if cdtype == torch.complex128 else
                   "; storage and transport complex64, all device arithmetic fp64 (parity bound 1e-3)"),
        }

    e2e = e2e_spectra(torch.complex128, "complex128")
    e2e_c64 = e2e_spectra(torch.complex64, "complex64")

    # ---- pinned-copy bandwidth per rank with every rank copying at once (what bounds e2e at N > 1) -----------
    def host_link():
        state = {}

        def prepare():
            if skip:
                raise skip
            n = 1 << 30
            (hb,) = arena.carve([((n,), torch.uint8)])
            state["h"], state["d"] = hb, torch.empty(n, dtype=torch.uint8, device=dev)
            state["d"].copy_(hb, non_blocking=True)
            torch.cuda.synchronize()

        _, err = guarded(prepare)
        ok, rk, msg = R.agree(err)
        if not ok:
            return {"error": msg, "failed_rank": rk}
        out = {}
        for name, fn in (("h2d", lambda: state["d"].copy_(state["h"], non_blocking=True)),
                         ("d2h", lambda: state["h"].copy_(state["d"], non_blocking=True))):
            dt, err = timed_wall(R, fn, 4)
            ok, rk, msg = R.agree(err)
            if not ok:
                return {"error": msg, "failed_rank": rk}
            gbs = 4 * state["h"].numel() / dt / 1e9
            out[name + "_gbs_per_rank_min"] = R.min(gbs)
            out[name + "_gbs_aggregate"] = R.sum(gbs)
        out["what"] = "1 GiB pinned <-> device copies, 4 back to back, all %d ranks at the same time" % world
        del state["d"]
        return out

    link = host_link()

    # ---- end-to-end from AUDIO: STFT and iSTFT on the device, only audio crosses PCIe ---
    def e2e_audio_leg():
        from overiva_b200 import stft as gstft
        from overiva_b200.synth import audio_batch_torch

        state = {}
        n_samples = int(DURATION_S * FS)

        def prepare():
            if skip:
                raise skip
            assert gstft.num_frames(n_samples, 4096, 2048) == T
            xh, yh = arena.carve([((B, n_samples, M), torch.float64), ((B, (T - 1) * 2048 + 4096, K), torch.float64)])
            for b0 in range(0, B, 128):
                nb = min(128, B - b0)
                xh[b0 : b0 + nb].copy_(audio_batch_torch(nb, n_samples, M, K, seed=99 + rank + 1000 * b0, device=dev))
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            state["xh"], state["yh"] = xh, yh
            for _ in range(2):
                gstft.separate_batch(xh, n_src=K, n_iter=N_ITER, framesize=4096, model=MODEL, out=yh)

        _, err = guarded(prepare)
        ok, rk, msg = R.agree(err)
        if not ok:
            return failed_leg(rk, msg)
        xh, yh = state["xh"], state["yh"]
        dt, err = timed_wall(R, lambda: gstft.separate_batch(xh, n_src=K, n_iter=N_ITER, framesize=4096, model=MODEL,
                                                             out=yh), e2e_steps)
        if err is None and not bool(torch.isfinite(yh).all()):
            err = FloatingPointError("non-finite output")
        ok, rk, msg = R.agree(err)
        if not ok:
            return failed_leg(rk, msg)
        dt = R.max(dt)
        return {
            "value": world * B * DURATION_S / (dt / e2e_steps), "unit": UNIT,
            "h2d_bytes_per_step": xh.numel() * 8, "d2h_bytes_per_step": yh.numel() * 8, "steps": e2e_steps,
            "ms_per_step": 1e3 * dt / e2e_steps,
            "api": "overiva_b200.stft.separate_batch(pinned host audio (B,N,M) float64, out=pinned): STFT 4096/2048 "
                   "-> loop -> iSTFT on the device, chunks of 32 mixtures on three streams",
        }

    del X, Y
    torch.cuda.empty_cache()
    e2e_audio = e2e_audio_leg()
    ob.clear_plan_cache()
    arena.release()
    torch.cuda.empty_cache()

    # ---- BASELINE configs 1-3: one short mixture each, device-resident, every rank on its own GPU -------------
    def configs_leg():
        from oracle import overiva_oracle as orc
        from overiva_b200.synth import convolutive_mixture, stft

        table = {
            "cfg1": (15.0, 4, dict(n_src=2, n_iter=20, model="laplace")),
            "cfg2": (15.0, 6, dict(n_iter=20, model="laplace")),
            "cfg3": (60.0, 8, dict(n_src=2, n_iter=20, model="gauss", init_eig=True)),
        }
        out = {}
        for name, (secs, m, kw) in table.items():
            if args.no_configs:
                raise RuntimeError("skipped (--no-configs)")
            mix, _ = convolutive_mixture(900 + len(name) + m, m, 2, duration=secs)
            Xn = stft(mix)
            Xd = torch.from_numpy(Xn).to(dev)
            for _ in range(3):
                Yd = ob.overiva(Xd, **kw)
            torch.cuda.synchronize()
            times = []
            for _ in range(9):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                Yd = ob.overiva(Xd, **kw)
                b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            ms_c = statistics.median(times)
            for _ in range(3):  # steady state of the drop-in call: staging buffers exist from the second call on
                Yn = ob.overiva(Xn, **kw)
            t0 = time.perf_counter()
            for _ in range(5):
                Yn = ob.overiva(Xn, **kw)
            ms_np = (time.perf_counter() - t0) / 5 * 1e3
            tt, ff, mm = Xn.shape
            kk = kw.get("n_src") or mm
            alg = (2 * kw["n_iter"] + 2) * ff * tt * mm * 16 + ff * tt * kk * 16
            rec = {"shape": [tt, ff, mm, kk], "ms_per_call": ms_c, "mixture_s_per_s": secs / (ms_c / 1e3),
                   "ms_per_call_numpy_in_out": ms_np, "hbm_floor_ms": alg / (peak * 1e9) * 1e3,
                   "frac_of_hbm_floor": alg / (peak * 1e9) * 1e3 / ms_c}
            if rank == 0 and not args.no_cpu:  # parity against the CPU oracle on the same input, in the same run
                t0 = time.perf_counter()
                Yo = orc.overiva(Xn, **kw)
                rec["cpu_oracle_s"] = time.perf_counter() - t0
                rec["rel_err_vs_oracle"] = float(np.linalg.norm(Yd.cpu().numpy() - Yo) / np.linalg.norm(Yo))
                rec["rel_err_numpy_path_vs_oracle"] = float(np.linalg.norm(Yn - Yo) / np.linalg.norm(Yo))
            out[name] = rec
            del Xd, Yd
        return out

    configs, err = guarded(configs_leg)
    ok, rk, msg = R.agree(err)
    if not ok:
        configs = {"error": msg, "failed_rank": rk}
    ob.clear_plan_cache()
    torch.cuda.empty_cache()

    # ---- BASELINE config 5: one long mixture, frequency-sharded, one K x T all-reduce per epoch ---------------
    cfg5 = cfg5_leg(args, R, torch, dist, ob, L, peak, fp64)

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n = 2 * cores
        kind = cpu_kind()
        pool = CpuPool(cores)
        v, wall, per = pool.throughput(n)
        pool.close()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d mixtures of the workload shape, %d worker processes x 1 BLAS thread, %.1f s wall, "
                         "%.2f s per mixture per core; %s" % (
                             n, cores, wall, statistics.median(per),
                             "the unmodified reference overiva.py (oracle/_ref bytecode)" if kind == "reference"
                             else "numpy port (reference bytecode absent)")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(B, world), "clocks": clocks,
            "e2e": e2e, "e2e_c64": e2e_c64, "e2e_audio": e2e_audio, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "fp64_peak": fp64, "host_link": link, "configs": configs,
            "cfg5_freq_sharded": cfg5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cfg5_leg(args, R, torch, dist, ob, L, hbm_peak, fp64):
    """Strong scaling of ONE mixture: bins sharded over the ranks (overiva_b200.distributed), the source-model
    statistic all-reduced once per epoch.  In-run parity: rank 0 also separates the WHOLE mixture on its own GPU and
    its shard of that result must equal the sharded run's to <= 1e-10."""
    from overiva_b200.distributed import overiva_freq_sharded, shard_bins
    from overiva_b200.synth import stft_domain_bins_torch

    rank, world, dev = R.rank, R.world, R.dev
    Tn = C5_T if args.cfg5_frames is None else args.cfg5_frames
    secs = C5_SECS * Tn / C5_T
    f0, f1 = shard_bins(C5_F, world, rank)
    state = {}

    def prepare():
        if args.no_cfg5:
            raise RuntimeError("skipped (--no-cfg5)")
        need = 3 * Tn * max(f1 - f0, 1) * C5_M * 16  # samples + grouped copy + output and covariances
        free, _ = torch.cuda.mem_get_info(dev)
        if need > free:
            raise MemoryError("cfg5 shard needs %.1f GB of device memory, %.1f GB free" % (need / 1e9, free / 1e9))
        state["X"] = stft_domain_bins_torch(Tn, C5_F, f0, f1, C5_M, C5_K, seed=4242, device=dev, n_interferers=C5_Q)

    _, err = guarded(prepare)
    ok, rk, msg = R.agree(err)
    if not ok:
        return failed_leg(rk, msg)
    Xl = state["X"]
    kw = dict(n_src=C5_K, n_iter=N_ITER, model="laplace")
    times, Yl, err = [], None, None
    for it in range(2 + 3):  # the collectives inside are issued by every rank in lock-step; local work is launches only
        R.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Yl = overiva_freq_sharded(Xl, C5_F, **kw)
        b.record()
        torch.cuda.synchronize()
        t = R.max(a.elapsed_time(b))
        if it >= 2:
            times.append(t)
    ms = statistics.median(times)
    # the all-reduce on its own: K x Tp doubles, 20 in a row
    Tp = (Tn + 31) // 32 * 32
    r2 = torch.zeros((1, C5_K, Tp), dtype=torch.float64, device=dev)
    ar_ms = None
    if world > 1:
        for _ in range(5):
            dist.all_reduce(r2)
        R.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(N_ITER):
            dist.all_reduce(r2)
        b.record()
        torch.cuda.synchronize()
        ar_ms = R.max(a.elapsed_time(b)) / N_ITER
    # in-run parity: the un-sharded separation on rank 0
    par = {}

    def check():
        if rank != 0 or args.no_cfg5_check:
            return
        Xfull = (stft_domain_bins_torch(Tn, C5_F, 0, C5_F, C5_M, C5_K, seed=4242, device=dev, n_interferers=C5_Q)
                 if world > 1 else Xl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ob.overiva(Xfull, **kw)
        a.record()
        Yfull = ob.overiva(Xfull, **kw)
        b.record()
        torch.cuda.synchronize()
        par["single_gpu_ms"] = a.elapsed_time(b)
        ref = Yfull[:, f0:f1]
        par["rel_err_shard_vs_single_gpu"] = float((torch.linalg.norm(Yl - ref) / torch.linalg.norm(ref)).item())
        par["finite"] = bool(torch.isfinite(torch.view_as_real(Yl)).all())

    _, err = guarded(check)
    ok, rk, msg = R.agree(err)
    flops = N_ITER * C5_F * Tn * (8 * C5_M * C5_K + 4 * C5_K * C5_M * (C5_M + 1)) + C5_F * Tn * (8 * C5_M * C5_K + 4 * C5_M * (C5_M + 1))
    alg_bytes = (2 * N_ITER + 2) * C5_F * Tn * C5_M * 16 + C5_F * Tn * C5_K * 16
    dfma = (fp64 or {}).get("dfma_tflops") or 37.0
    t_hbm = alg_bytes / (hbm_peak * 1e9) * 1e3 / world
    t_fma = flops / (dfma * 1e12) * 1e3 / world
    out = {
        "value": secs / (ms / 1e3), "unit": UNIT, "ms_per_call": ms, "mixture_s_per_s": secs / (ms / 1e3),
        "n_gpus": world, "scaling": "strong", "shape": [Tn, C5_F, C5_M, C5_K], "bins_rank0": f1 - f0,
        "algorithmic_flops": flops, "algorithmic_bytes": alg_bytes,
        "bound_ms": {"hbm": t_hbm, "fma": t_fma, "fma_peak_tflops": dfma,
                     "fma_peak_source": "oiva_fp64_peak(DFMA) measured in this run" if (fp64 or {}).get("dfma_tflops")
                     else "nominal 37 TFLOP/s"},
        "frac_of_max_hbm_fma": max(t_hbm, t_fma) / ms,
        "allreduce_bytes": C5_K * Tp * 8, "allreduce_ms": ar_ms,
        "collective": "torch.distributed all_reduce(SUM) over NCCL of the (K, Tp) float64 statistic, once per epoch",
    }
    out.update(par)
    if not ok:
        out["check_error"] = msg
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="mixtures per GPU (default 512 = BASELINE cfg4 / 8)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the oracle checks of `configs`")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE cfg1-3 leg")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the frequency-sharded cfg5 leg")
    ap.add_argument("--no-cfg5-check", action="store_true", help="skip the single-GPU parity run of the cfg5 leg")
    ap.add_argument("--cfg5-frames", type=int, default=None, help="frames of the cfg5 mixture (default 14061)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
