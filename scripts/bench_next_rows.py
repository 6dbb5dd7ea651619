#!/usr/bin/env python
"""Measurements for the SURVEY 8(f) rows 2 and 3: the batched sweep driver and the device-side convergence monitor.
One JSON line each."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import overiva_b200 as ob  # noqa: E402
from overiva_b200 import metrics, monitor, stft, sweep  # noqa: E402
from overiva_b200.synth import convolutive_mixture  # noqa: E402

torch.cuda.set_device(0)

# ---- row 2: sweep (reference: one CPU task per mixture and algorithm, overiva_sim.py) -------------------------
params = dict(sweep.DEFAULT_PARAMETERS, n_repeat=8, n_targets_list=[1, 2], n_mics_list=[2, 4, 6], duration=10.0)
t0 = time.perf_counter()
rows = []
segs = sweep.run(params, batch=64, progress=lambda *a: rows.append(a))
wall = time.perf_counter() - t0
n_mix = len(segs)
gpu_s = sum(r[3] * r[4] for r in rows)  # batch size x seconds per mixture, per (algorithm, shape) call
summ = sweep.summarise(segs, params["fs"])
print(json.dumps({
    "row": "sweep", "mixtures": n_mix, "records": sum(len(s) for s in segs), "audio_s_per_mixture": params["duration"],
    "wall_s_total_incl_host_mixture_generation_and_metrics": wall, "gpu_algorithm_s_total": gpu_s,
    "algorithm_runs": int(sum(r[3] for r in rows)),
    "mean_sir_improvement_db": {"%s/%d/%d" % (r["algorithm"], r["n_targets"], r["n_mics"]): round(r["sir_improvement"], 2)
                                for r in summ},
    "real_time_factor": {"%s/%d/%d" % (r["algorithm"], r["n_targets"], r["n_mics"]): r["runtime_per_s"] for r in summ},
}), flush=True)

# ---- row 3: convergence monitor -------------------------------------------------------------------------------
mix, images = convolutive_mixture(7, 4, 2, duration=15.0)
L_, hop = 4096, 2048
Xd = stft.analysis(torch.from_numpy(mix).cuda(), L_, hop, win=stft.hann(L_), pad_front=L_ - hop)
ws = stft.compute_synthesis_window(stft.hann(L_), hop)


def timed(cb_factory, reps=5):
    best = 1e30
    for _ in range(reps):
        cb = cb_factory()
        torch.cuda.synchronize()
        t = time.perf_counter()
        ob.overiva(Xd, n_src=2, n_iter=100, callback=cb)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, cb


def host_callback():
    from overiva_b200.synth import istft  # numpy iSTFT (rectangular synthesis): the host-side equivalent

    SDR = []

    def cb(Y):
        y = istft(Y.cpu().numpy(), L_, hop)
        y = y[:, np.argsort(np.std(y, axis=0))[::-1]]
        m = min(y.shape[0] - (L_ - hop), images.shape[1])
        sdr, sir, _ = metrics.bss_eval(images[:, :m, 0], y[L_ - hop : L_ - hop + m, :2].T)
        SDR.append(sdr)

    cb.SDR = SDR
    return cb


ob.overiva(Xd, n_src=2, n_iter=3)
t_none, _ = timed(lambda: None)
t_dev, mon = timed(lambda: monitor.ConvergenceMonitor(images, framesize=L_, delay=L_ - hop))
t_host, hcb = timed(host_callback, reps=2)
print(json.dumps({
    "row": "monitor", "workload": "one 15 s mixture, M=4 K=2, n_iter=100 -> 10 callbacks",
    "ms_no_callback": t_none * 1e3, "ms_device_monitor": t_dev * 1e3, "ms_host_callback": t_host * 1e3,
    "ms_per_callback_device": (t_dev - t_none) * 100, "ms_per_callback_host": (t_host - t_none) * 100,
    "final_sdr_device": [float(v) for v in mon.SDR[-1]], "calls": len(mon.SDR),
}), flush=True)

# ---- row 3b: the same monitor with mir_eval's 512-tap distortion filters (cross-correlations on the device) ----------
t_dev512, mon512 = timed(lambda: monitor.ConvergenceMonitor(images, framesize=L_, delay=L_ - hop, filter_length=512), reps=2)


def host_callback_512():
    from overiva_b200.synth import istft

    SDR = []

    def cb(Y):
        y = istft(Y.cpu().numpy(), L_, hop)
        y = y[:, np.argsort(np.std(y, axis=0))[::-1]]
        m = min(y.shape[0] - (L_ - hop), images.shape[1])
        sdr, sir, _, _ = metrics.bss_eval_sources(images[:2, :m, 0], y[L_ - hop : L_ - hop + m, :2].T, flen=512)
        SDR.append(sdr)

    cb.SDR = SDR
    return cb


t0 = time.perf_counter()
hcb512 = host_callback_512()
ob.overiva(Xd, n_src=2, n_iter=10, callback=hcb512)  # one callback (epoch 0) is enough to time the host metric
torch.cuda.synchronize()
t_host512_one = time.perf_counter() - t0
print(json.dumps({
    "row": "monitor_512tap", "workload": "one 15 s mixture, M=4 K=2, n_iter=100 -> 10 callbacks, bss_eval_sources with 512 taps",
    "ms_device_monitor_512": t_dev512 * 1e3, "ms_per_callback_device_512": (t_dev512 - t_none) * 100,
    "ms_per_callback_host_512": t_host512_one * 1e3, "final_sdr_device_512": [float(v) for v in mon512.SDR[-1]],
    "first_sdr_host_512": [float(v) for v in hcb512.SDR[0]], "first_sdr_device_512": [float(v) for v in mon512.SDR[0]],
}), flush=True)
