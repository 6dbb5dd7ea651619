#!/usr/bin/env python
"""Audio in -> audio out: device time of the STFT / iSTFT kernels at the bench shape and the
end-to-end rate of ``overiva_b200.stft.separate_batch`` from pinned host audio, next to the spectra-in / spectra-out
path bench.py reports as ``e2e``.  One JSON line.

    python scripts/bench_audio.py [--mixtures 512] [--reps 3]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from overiva_b200 import _lib as L  # noqa: E402
from overiva_b200 import core, stft  # noqa: E402


def ev_time(fn, reps):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mixtures", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--mics", type=int, default=6)
    ap.add_argument("--src", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, M, K, fs, Lf, hop = args.mixtures, args.mics, args.src, 16000, 4096, 2048
    N = int(args.seconds * fs)
    T = stft.num_frames(N, Lf, hop)
    F = Lf // 2 + 1
    lib = L.load()
    from overiva_b200.synth import audio_batch_torch

    g = torch.Generator(device=dev)
    g.manual_seed(1)
    x = audio_batch_torch(B, N, M, K, seed=4321, device=dev)
    xh = torch.empty((B, N, M), dtype=torch.float64).pin_memory()
    xh.copy_(x)
    res = {"workload": "%d mixtures x %.0f s @ 16 kHz, M=%d K=%d, framesize 4096 hop 2048 -> (T=%d, F=%d), 20 iterations"
                       % (B, args.seconds, M, K, T, F)}
    # --- kernel times on the device
    plan = core.DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    wa = torch.from_numpy(stft.hann(Lf)).to(dev)
    ws = torch.from_numpy(stft.compute_synthesis_window(stft.hann(Lf), hop)).to(dev)
    tw = stft._twiddles(Lf, dev)
    st = core._stream_ptr(dev)

    def ana():
        L.check(lib.oiva_stft_analysis(core._ptr(x), 0, N * M, M, 1, N, 0, core._ptr(wa), core._ptr(tw),
                                       C.c_void_p(plan.samples_ptr), 1, B, T, M, Lf, hop, L.C128, st), "analysis")

    res["analysis_grouped_ms"] = ev_time(ana, args.reps)
    X = torch.empty((B, T, F, M), dtype=torch.complex128, device=dev)

    def ana_plain():
        L.check(lib.oiva_stft_analysis(core._ptr(x), 0, N * M, M, 1, N, 0, core._ptr(wa), core._ptr(tw),
                                       core._ptr(X), 0, B, T, M, Lf, hop, L.C128, st), "analysis")

    res["analysis_plain_ms"] = ev_time(ana_plain, args.reps)
    res["relayout_plus_cov_ms"] = ev_time(lambda: plan.load(X), args.reps)
    res["adopt_cov_ms"] = ev_time(lambda: plan.adopt_samples(), args.reps)
    del X
    Y = torch.randn((B, T, F, K), generator=g, device=dev, dtype=torch.float64).to(torch.complex128)
    res["synthesis_ms"] = ev_time(lambda: stft._synthesis_dev(Y, Lf, hop, ws), args.reps)
    del Y, plan
    torch.cuda.empty_cache()
    # --- device-resident audio -> audio
    def dev_call():
        return stft.separate(x, n_src=K, n_iter=20, framesize=Lf)

    ms = ev_time(dev_call, args.reps)
    res["device_resident_ms"] = ms
    res["device_resident_mixture_s_per_s"] = B * args.seconds / (ms * 1e-3)
    del x
    torch.cuda.empty_cache()
    # --- end to end from pinned host audio
    n_out = (T - 1) * hop + Lf
    yh = torch.empty((B, n_out, K), dtype=torch.float64).pin_memory()
    stft.separate_batch(xh, n_src=K, n_iter=20, framesize=Lf, out=yh)
    best = 1e30
    for _ in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stft.separate_batch(xh, n_src=K, n_iter=20, framesize=Lf, out=yh)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    res["e2e_audio"] = {"value": B * args.seconds / best, "unit": "mixture-s/s", "ms": best * 1e3,
                        "h2d_bytes_per_step": xh.numel() * 8, "d2h_bytes_per_step": yh.numel() * 8,
                        "api": "overiva_b200.stft.separate_batch(pinned host audio, out=pinned): chunks of 32"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
