#!/usr/bin/env python
"""Where an epoch of the resident loop goes: runs the short BASELINE shapes once with OIVA_RES_TRACE set (clock64 stamps
of thread 0 of every CTA at the phase boundaries, csrc/resident.cu) and prints the median SM cycles per phase over the
CTAs and the epochs >= 5, owners (first CTA of a cluster) and the other CTAs separately."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from overiva_b200 import _lib as L
from overiva_b200.core import DemixPlan
from overiva_b200.synth import stft_domain_batch_torch

SHAPES = {"cfg1": (1, 116, 2049, 4, 2), "cfg2": (1, 116, 2049, 6, 6)}
PHASES = ["1 demix+power", "grid barrier 1", "2 sum over groups", "grid barrier 2", "3 gamma, phi", "4 covariance",
          "hand-over 1", "5 sum + sweep (owner)", "hand-over 2"]
dev = torch.device("cuda", 0)
for name, (B, T, F, M, K) in SHAPES.items():
    X = stft_domain_batch_torch(B, T, F, M, K, seed=3, device=dev, chunk=1)
    plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    plan.load(X); plan.init(L.INIT_EYE); plan.iterate(20); torch.cuda.synchronize()
    path = "/tmp/res_trace_%s.txt" % name
    os.environ["OIVA_RES_TRACE"] = path
    plan.init(L.INIT_EYE); plan.iterate(20); torch.cuda.synchronize()
    del os.environ["OIVA_RES_TRACE"]
    a = np.loadtxt(path, dtype=np.int64)
    cta, ep, st = a[:, 0], a[:, 1], a[:, 2:]
    grid = int(cta.max()) + 1
    SG = grid // (B * ((F + 31) // 32))
    sub = st[:, 10:]  # inside phase 5 (owner only): after the sum of source 0 | (determined: W rescale) | source 0 | ...
    st = st[:, :10]
    d = np.diff(st, axis=1)
    keep = ep >= 5
    out = {"config": name, "grid": grid, "slices_per_group": SG, "unit": "SM cycles (median over CTAs and epochs 5..19)"}
    for who, sel in (("owner", (cta % SG) == 0), ("other", (cta % SG) != 0)):
        m = keep & sel
        if m.any():
            out[who] = {PHASES[j]: float(np.median(d[m, j])) for j in range(len(PHASES))}
            out[who]["epoch"] = float(np.median(st[m, -1] - st[m, 0]))
            if who == "owner":
                pts = np.concatenate([st[m, 7:8], sub[m]], axis=1)
                names = (["sum of source 0", "W rescale + pair sweep of source 0 (|| sum of source 1)", "pair sweep of source 1", "pair sweep of source 2"] if M == K and M >= 3 else
                         ["sum of source 0", "sweep of source 0 (|| sum of source 1)", "sweep of source 1", "sweep of source 2"])
                out["inside phase 5"] = {names[j]: float(np.median(pts[:, j + 1] - pts[:, j])) for j in range(4)
                                         if np.all(pts[:, j + 1] > 0)}
    print(json.dumps(out), flush=True)
    del plan, X
