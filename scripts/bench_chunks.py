#!/usr/bin/env python
"""Chunk-size sweep of the host pipelines (overiva_batch from pinned spectra, separate_batch from pinned audio)
at the bench workload: the pipelines overlap H2D / compute / D2H, the un-overlapped head and tail shrink with the
chunk.  One JSON line per (path, chunk)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from overiva_b200 import stft  # noqa: E402
from overiva_b200.synth import audio_batch_torch, stft_domain_batch_torch  # noqa: E402

B, T, F, M, K, N = 512, 116, 2049, 6, 2, 240000
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def timeit(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


Xh = torch.empty((B, T, F, M), dtype=torch.complex128, pin_memory=True)
Xh.copy_(stft_domain_batch_torch(B, T, F, M, K, seed=1234, device=dev))
Yh = torch.empty((B, T, F, K), dtype=torch.complex128, pin_memory=True)
torch.cuda.empty_cache()
for chunk in (16, 32, 64, 128):
    dt = timeit(lambda: ob.overiva_batch(Xh, n_src=K, n_iter=20, out=Yh, chunk=chunk))
    print(json.dumps({"path": "spectra", "chunk": chunk, "ms": dt * 1e3, "mixture_s_per_s": B * 15.0 / dt,
                      "h2d_GBps": Xh.numel() * 16 / dt / 1e9}), flush=True)
del Xh, Yh
xh = torch.empty((B, N, M), dtype=torch.float64, pin_memory=True)
xh.copy_(audio_batch_torch(B, N, M, K, seed=99, device=dev))
yh = torch.empty((B, (T - 1) * 2048 + 4096, K), dtype=torch.float64, pin_memory=True)
torch.cuda.empty_cache()
for chunk in (16, 32, 64, 128):
    dt = timeit(lambda: stft.separate_batch(xh, n_src=K, n_iter=20, out=yh, chunk=chunk))
    print(json.dumps({"path": "audio", "chunk": chunk, "ms": dt * 1e3, "mixture_s_per_s": B * 15.0 / dt,
                      "h2d_GBps": xh.numel() * 8 / dt / 1e9}), flush=True)
