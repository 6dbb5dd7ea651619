#!/usr/bin/env python
"""OGIVE (ive.py) on BASELINE config 1's mixture shape, default arguments (n_iter=4000, tol=1e-3): device time and the
numpy oracle beside it."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from overiva_b200.synth import convolutive_mixture, stft  # noqa: E402

mix, _ = convolutive_mixture(905, 4, 1, duration=15.0)
X = stft(mix)
Xd = torch.from_numpy(X).cuda()
n_calls = []
ob.ogive(Xd, n_iter=50)
torch.cuda.synchronize()
res = {}
for upd in ("demix", "switching"):
    t0 = time.perf_counter()
    Y, w = ob.ogive(Xd, update=upd, return_filters=True)
    torch.cuda.synchronize()
    res[upd] = {"gpu_s": time.perf_counter() - t0}
if "--cpu" in sys.argv:
    from oracle import overiva_oracle as orc

    t0 = time.perf_counter()
    Yo, wo = orc.ogive(X, update="switching", return_filters=True)
    res["switching"]["cpu_oracle_s"] = time.perf_counter() - t0
    res["switching"]["rel_err_w"] = float(np.linalg.norm(w.cpu().numpy() - wo) / np.linalg.norm(wo))
print(json.dumps({"shape": list(X.shape), "ogive_default_args": res}))
