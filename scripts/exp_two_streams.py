#!/usr/bin/env python
"""Experiment: the 512-mixture step as S independent sub-batches on S streams (mixtures are independent, so the
latency-bound sweep of one sub-batch can overlap the streaming kernels of another)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from overiva_b200 import _lib as L  # noqa: E402
from overiva_b200.core import DemixPlan  # noqa: E402
from overiva_b200.synth import stft_domain_batch_torch  # noqa: E402

B, T, F, M, K = 512, 116, 2049, 6, 2
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
X = stft_domain_batch_torch(B, T, F, M, K, seed=1234, device=dev)
Y = torch.empty((B, T, F, K), dtype=torch.complex128, device=dev)
for S in (1, 2, 4):
    nb = B // S
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    plans = []
    for i in range(S):
        with torch.cuda.stream(streams[i]):
            plans.append(DemixPlan(nb, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev))
    torch.cuda.synchronize()

    def step():
        main = torch.cuda.current_stream(dev)
        for i in range(S):
            streams[i].wait_stream(main)
            with torch.cuda.stream(streams[i]):
                p = plans[i]
                p.load(X[i * nb : (i + 1) * nb])
                p.init(L.INIT_EYE)
                p.iterate(20)
                p.output(True, out=Y[i * nb : (i + 1) * nb])
        for i in range(S):
            main.wait_stream(streams[i])

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    for p in plans:
        p.raise_on_failure()
    print(json.dumps({"streams": S, "ms_per_step": ms, "mixture_s_per_s": B * 15.0 / ms * 1e3}), flush=True)
    del plans
    torch.cuda.empty_cache()
