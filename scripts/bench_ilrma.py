#!/usr/bin/env python
"""ILRMA (overiva_b200.ilrma, the stand-in for pra.bss.ilrma of overiva_oneshot.py:331-339) on one GPU against its numpy
restatement on one host core: time per call and agreement, for the demo driver's shape (15 s @ 16 kHz, STFT 4096/2048)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from oracle import ilrma_oracle as ilo  # noqa: E402
from overiva_b200.synth import convolutive_mixture, stft  # noqa: E402

for M in (2, 4, 6):
    mix, _ = convolutive_mixture(500 + M, M, 2, duration=15.0)
    X = stft(mix, 4096, 2048)
    T, F, _ = X.shape
    rng = np.random.default_rng(M)
    T0, V0 = 0.1 + 0.9 * rng.random((M, F, 2)), 0.1 + 0.9 * rng.random((M, T, 2))
    Xd = torch.from_numpy(X).cuda()
    kw = dict(n_iter=20, n_components=2, proj_back=True, T0=T0, V0=V0)
    for _ in range(2):
        Y = ob.ilrma(Xd, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        Y = ob.ilrma(Xd, **kw)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    Yo = ilo.ilrma(X, **kw)
    cpu_s = time.perf_counter() - t0
    err = float(np.linalg.norm(Y.cpu().numpy() - Yo) / np.linalg.norm(Yo))
    print(json.dumps({"algo": "ilrma", "shape": [T, F, M, M], "n_components": 2, "n_iter": 20, "ms_per_call_gpu": gpu_ms,
                      "s_per_call_numpy_1core": cpu_s, "speedup": cpu_s * 1e3 / gpu_ms, "rel_err_vs_oracle": err}), flush=True)
