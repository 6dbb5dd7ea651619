#!/usr/bin/env python
"""Device time of the IP sweep (oiva_ip_update) for one (M, K): thread-per-bin vs row-owner kernel
(OIVA_SOLVER_ROWOWNER=1 selects the latter at process start).  python scripts/bench_solver.py M K [B [T [F]]]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from overiva_b200 import _lib as L  # noqa: E402
from overiva_b200 import core  # noqa: E402
from overiva_b200.synth import stft_domain_batch_torch  # noqa: E402

M, K = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
T = int(sys.argv[4]) if len(sys.argv) > 4 else 116
F = int(sys.argv[5]) if len(sys.argv) > 5 else 2049
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
X = stft_domain_batch_torch(B, T, F, M, K, seed=5, device=dev, chunk=1 if T > 2000 else 32)
plan = core.DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
plan.load(X)
plan.init(L.INIT_EYE)
plan.enable_timing(True)
plan.iterate(3)
torch.cuda.synchronize()
plan.read_timing()
plan.iterate(10)
torch.cuda.synchronize()
t = plan.read_timing()
plan.raise_on_failure()
print(json.dumps({"M": M, "K": K, "B": B, "T": T, "F": F, "rowowner": os.environ.get("OIVA_SOLVER_ROWOWNER", "0"),
                  "ms_per_launch": {k: v[0] / max(v[1], 1) for k, v in t.items()}}))
