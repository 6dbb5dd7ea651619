#!/usr/bin/env python
"""A/B of the resident loop's synchronisation modes: 20-epoch loop time of the short BASELINE shapes, modes interleaved,
median of 15; also checks the demixing matrices are bit-identical across modes.  Modes (OIVA_RES_CLUSTER, OIVA_RES_POLL,
OIVA_RES_TAGGED): "f0" flag hand-over inside a bin group + acquire polls, "f2" relaxed polls + one acquire load, "c2" the
slices of a bin group as a thread-block cluster (cluster barriers instead of the arrive counter and the flag), "f2t" /
"c2t" the same with the statistic words tagged by the epoch's parity in their sign bit instead of two grid barriers per
epoch, "c2t+inv" with the tracked-inverse determined sweep (OIVA_RES_TRACKED; not bit-identical: another factorisation)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from overiva_b200 import _lib as L
from overiva_b200.core import DemixPlan
from overiva_b200.synth import stft_domain_batch_torch

SHAPES = {"cfg1": (1, 116, 2049, 4, 2), "cfg2": (1, 116, 2049, 6, 6), "cfg3": (1, 467, 2049, 8, 2)}
dev = torch.device("cuda", 0)
for name, (B, T, F, M, K) in SHAPES.items():
    X = stft_domain_batch_torch(B, T, F, M, K, seed=3, device=dev, chunk=1)
    plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, torch.complex128, dev)
    plan.load(X)
    MODES = {"f0": ("0", "0", "0", "0"), "c2": ("1", "2", "0", "0"), "f2t": ("0", "2", "1", "0"), "c2t": ("1", "2", "1", "0"),
             "c2t+inv": ("1", "2", "1", "1")}
    times = {m: [] for m in MODES}
    Ws = {}
    for rep in range(16):
        for mode, (cl, po, ta, tr) in MODES.items():
            os.environ["OIVA_RES_CLUSTER"], os.environ["OIVA_RES_POLL"], os.environ["OIVA_RES_TAGGED"] = cl, po, ta
            os.environ["OIVA_RES_TRACKED"] = tr
            plan.init(L.INIT_EYE); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.iterate(20); e1.record(); torch.cuda.synchronize()
            if rep: times[mode].append(e0.elapsed_time(e1))
            Ws[mode] = plan.filters().clone()
    same = {m: bool(torch.equal(Ws["f0"], Ws[m])) for m in MODES}
    diff = {m: float((Ws["f0"] - Ws[m]).abs().max() / Ws["f0"].abs().max()) for m in MODES}
    print(json.dumps({"config": name, "loop_ms_median": {m: sorted(v)[len(v) // 2] for m, v in times.items()},
                      "loop_ms_min": {m: min(v) for m, v in times.items()}, "bit_identical_to_f0": same, "max_rel_diff_to_f0": diff,
                      "status": plan.status()}), flush=True)
    del plan, X
