#!/usr/bin/env python
"""Quick check of the persistent single-launch loop (csrc/resident.cuh) against the kernel-per-step loop and the numpy
oracle, plus device-resident timings of BASELINE configs 1-2; exits non-zero on mismatch.  Run under `timeout`."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from oracle import overiva_oracle as orc  # noqa: E402
from overiva_b200 import _lib as L  # noqa: E402
from overiva_b200.core import DemixPlan  # noqa: E402
from overiva_b200.synth import convolutive_mixture, small_test_mixture, stft  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_plan(Xd, K, model, n_iter, resident):
    if resident:
        os.environ.pop("OIVA_NO_RESIDENT", None)
    else:
        os.environ["OIVA_NO_RESIDENT"] = "1"
    B, T, F, M = Xd.shape
    plan = DemixPlan(B, T, F, M, K, model, Xd.dtype, Xd.device)
    plan.load(Xd)
    plan.init(L.INIT_EYE)
    l0 = plan.launches
    plan.iterate(n_iter)
    nl = plan.launches - l0
    Y = plan.output(True).cpu().numpy()
    W = plan.filters().cpu().numpy()
    st = plan.status()
    return Y, W, nl, st


worst = 0.0
for M, K, n_samples, frame in [(4, 2, 3000, 64), (2, 1, 900, 32), (6, 6, 20000, 512), (3, 3, 5000, 128), (8, 4, 16000, 256),
                               (5, 2, 60000, 2048), (7, 3, 4000, 64), (6, 2, 9000, 256)]:
    for B in (1, 2):
        X = np.stack([small_test_mixture(70 + b, M, min(K, 2), n_samples=n_samples, frame=frame, hop=frame // 2)
                      for b in range(B)])
        Xd = torch.from_numpy(X).cuda()
        Yr, Wr, nl, st = run_plan(Xd, K, L.MODEL_LAPLACE, 10, True)
        Yk, Wk, nk, _ = run_plan(Xd, K, L.MODEL_LAPLACE, 10, False)
        e = max(rel(Yr, Yk), rel(Wr, Wk))
        worst = max(worst, e)
        print("M=%d K=%d B=%d T=%d F=%d: resident launches %d (kernel loop %d) status %d  rel err vs kernel loop %.2e"
              % (M, K, B, X.shape[1], X.shape[2], nl, nk, st, e), flush=True)
        assert nl == 1 and st == 0 and e < 1e-11, (nl, st, e)
os.environ.pop("OIVA_NO_RESIDENT", None)
for name, m, secs, kw in [("cfg1", 4, 15.0, dict(n_src=2, n_iter=20, model="laplace")),
                          ("cfg2", 6, 15.0, dict(n_iter=20, model="laplace")),
                          ("cfg3", 8, 60.0, dict(n_src=2, n_iter=20, model="gauss", init_eig=True))]:
    mix, _ = convolutive_mixture(900 + m, m, 2, duration=secs)
    Xn = stft(mix)
    Xd = torch.from_numpy(Xn).cuda()
    Yo = orc.overiva(Xn, **kw)
    for mode in ("resident", "kernel_loop"):
        if mode == "kernel_loop":
            os.environ["OIVA_NO_RESIDENT"] = "1"
        else:
            os.environ.pop("OIVA_NO_RESIDENT", None)
        ob.clear_plan_cache()
        for _ in range(3):
            Y = ob.overiva(Xd, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(9):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            Y = ob.overiva(Xd, **kw)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        # the epoch loop alone
        T_, F_, M_ = Xn.shape
        K_ = kw.get("n_src") or M_
        plan = DemixPlan(1, T_, F_, M_, K_, L.MODEL_GAUSS if kw["model"] == "gauss" else L.MODEL_LAPLACE, Xd.dtype, Xd.device)
        plan.load(Xd[None])
        plan.init(L.INIT_EYE)
        plan.iterate(20)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            plan.iterate(20)
        b.record()
        torch.cuda.synchronize()
        print("%s %s: %.3f ms per call (median of 9), loop of 20 epochs alone %.3f ms, rel err vs oracle %.2e"
              % (name, mode, sorted(ts)[4], a.elapsed_time(b) / 5, rel(Y.cpu().numpy(), Yo)), flush=True)
        assert rel(Y.cpu().numpy(), Yo) < 1e-10
print("RESIDENT_OK")
