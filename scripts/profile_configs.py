#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown (cov / power / solve) of the loop for the BASELINE shapes, one GPU."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from overiva_b200 import _lib as L
from overiva_b200.core import DemixPlan
from overiva_b200.synth import stft_domain_batch_torch

SHAPES = {"cfg1": (1, 116, 2049, 4, 2), "cfg2": (1, 116, 2049, 6, 6), "cfg3": (1, 467, 2049, 8, 2),
          "cfg5": (1, 14061, 2049, 16, 4), "cfg5_shard8": (1, 14061, 256, 16, 4), "cfg4_b64": (64, 116, 2049, 6, 2),
          # determined AuxIVA batches of the paper sweep
          "det4_b256": (256, 116, 2049, 4, 4), "det6_b256": (256, 116, 2049, 6, 6), "det8_b256": (256, 116, 2049, 8, 8)}
dev = torch.device("cuda", 0)
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else SHAPES):
    B, T, F, M, K = SHAPES[name]
    X = stft_domain_batch_torch(B, T, F, M, K, seed=3, device=dev, chunk=1)
    plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    plan.load(X); plan.init(L.INIT_EYE); plan.iterate(3); torch.cuda.synchronize()
    plan.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.iterate(10); e1.record(); torch.cuda.synchronize()
    tm = plan.read_timing()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(); plan.load(X); e3.record(); torch.cuda.synchronize()
    Y = torch.empty((B, T, F, K), dtype=torch.complex128, device=dev)
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.init(L.INIT_EYE); torch.cuda.synchronize()
    e4.record(); plan.output(True, out=Y); e5.record(); torch.cuda.synchronize()
    xb = B * T * F * M * 16
    print(json.dumps({"config": name, "B,T,F,M,K": [B, T, F, M, K], "x_GB": xb / 1e9,
                      "ms_per_epoch": e0.elapsed_time(e1) / 10,
                      "cov_ms": tm["cov"][0] / 10, "power_ms": tm["power"][0] / 10, "solve_ms": tm["solve"][0] / 10,
                      "cov_GBps": xb / (tm["cov"][0] / 10 * 1e-3) / 1e9, "power_GBps": xb / (tm["power"][0] / 10 * 1e-3) / 1e9,
                      "load_ms": e2.elapsed_time(e3), "output_ms": e4.elapsed_time(e5), "status": plan.status()}), flush=True)
    del plan, X, Y
    torch.cuda.empty_cache()
