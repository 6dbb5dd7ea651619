#!/usr/bin/env python
"""Device-resident timings of ALL five BASELINE.json configurations on one GPU (bench.py measures the headline
configuration 4 with the full contract; this script is the supplementary table).  One JSON line per config.

    python scripts/bench_configs.py [--configs cfg1,cfg2,cfg3,cfg5] [--cpu]

cfg5 (one 10-minute 48 kHz mixture, M=16, K=4) runs un-sharded here; its 8-way frequency-sharded form is
scripts/bench_freq_sharded.py under torchrun.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from overiva_b200.synth import convolutive_mixture, stft, stft_domain_batch_torch  # noqa: E402

CONFIGS = {
    # name: (audio seconds, T, F, M, kwargs, generator)
    "cfg1": (15.0, 116, 2049, 4, dict(n_src=2, n_iter=20, model="laplace"), "conv"),
    "cfg2": (15.0, 116, 2049, 6, dict(n_iter=20, model="laplace"), "conv"),
    "cfg3": (60.0, 467, 2049, 8, dict(n_src=2, n_iter=20, model="gauss", init_eig=True), "conv"),
    "cfg5": (600.0, 14061, 2049, 16, dict(n_src=4, n_iter=20, model="laplace"), "stft"),
}


def make_input(name, dev):
    secs, T, F, M, kw, gen = CONFIGS[name]
    if gen == "conv":
        mix, _ = convolutive_mixture(900 + len(name), M, 2, duration=secs)
        X = stft(mix)
        assert X.shape == (T, F, M), X.shape
        return torch.from_numpy(X).to(dev), X
    K = kw.get("n_src") or M
    Xd = stft_domain_batch_torch(1, T, F, M, K, seed=77, device=dev, chunk=1)[0]
    return Xd, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg5")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--cpu", action="store_true", help="also time the numpy oracle (single process) on cfg1-3")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = 6548.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        pass
    for name in args.configs.split(","):
        secs, T, F, M, kw, gen = CONFIGS[name]
        K = kw.get("n_src") or M
        Xd, Xnp = make_input(name, dev)
        for _ in range(2):
            Y = ob.overiva(Xd, **kw)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            Y = ob.overiva(Xd, **kw)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        assert bool(torch.isfinite(Y.real).all())
        ms = statistics.median(times)
        n_iter = kw["n_iter"]
        alg = ((2 * n_iter + 2) * F * T * M * 16 + F * T * K * 16)
        line = {
            "config": name, "shape": {"T": T, "F": F, "M": M, "K": K}, "kwargs": kw, "audio_s": secs,
            "ms_per_call_device_resident": ms, "mixture_s_per_s": secs / (ms / 1e3),
            "algorithmic_GB": alg / 1e9, "algorithmic_GBps": alg / (ms / 1e3) / 1e9,
            "frac_of_hbm_peak": alg / (ms / 1e3) / 1e9 / peak,
            "x_bytes_MB": T * F * M * 16 / 1e6, "data": "synthetic (%s)" % gen,
            "note": "X (and its grouped copy) fit the 126 MB L2" if 2 * T * F * M * 16 < 120e6 else "",
        }
        if Xnp is not None:  # numpy in -> numpy out through the drop-in entry point (PCIe both ways)
            t0 = time.perf_counter()
            for _ in range(3):
                ob.overiva(Xnp, **kw)
            line["ms_per_call_numpy_in_out"] = (time.perf_counter() - t0) / 3 * 1e3
        if args.cpu and Xnp is not None:
            from oracle import overiva_oracle as orc

            t0 = time.perf_counter()
            Yo = orc.overiva(Xnp, **kw)
            line["cpu_oracle_s"] = time.perf_counter() - t0
            line["cpu_oracle_mixture_s_per_s"] = secs / line["cpu_oracle_s"]
            line["rel_err_Y_vs_oracle"] = float(np.linalg.norm(Y.cpu().numpy() - Yo) / np.linalg.norm(Yo))
        print(json.dumps(line), flush=True)
        del Xd, Y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
