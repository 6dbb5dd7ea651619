"""Batched sweep driver: the role of the reference's ``overiva_sim.py`` + ``rrtools`` task farm.

The reference fans one ``one_loop(args)`` call per simulated room out to an ipyparallel cluster
(``rrtools/dumbparallel.py:198-334``); every call simulates a mixture, runs each configured algorithm on it on one CPU
core and appends one record per algorithm to ``data.json`` (``overiva_sim.py:245-350``).  Here the same argument
list is grouped by mixture shape ``(n_targets, n_mics)`` and each group goes through the GPU as ONE batch per
algorithm configuration: analysis -> ``overiva_batch`` -> synthesis stay on the device, only audio crosses PCIe.

Kept from the reference:
  * the argument enumeration order and per-simulation seeds (``overiva_sim.py:356-392``),
  * which algorithm runs for which case (``overiva_sim.py:250-256``: no ``auxiva_pca`` for one target, ``ogive`` only
    for one target; unknown algorithms are skipped, ``:316-317``),
  * the evaluation rule of ``convergence_callback`` (``overiva_sim.py:210-232``): synthesis, reorder by decreasing
    power unless the algorithm's base name is in ``overdet_algos`` (the reference tests ``name``, not ``full_name``), drop the ``framesize // 2`` delay of the STFT state buffer,
    score the first ``n_targets`` outputs plus a noise channel against targets + background,
  * the record schema (``overiva_sim.py:258-271``) and the ``data.json`` / ``parameters.json`` / ``arguments.json``
    files ``overiva_sim_plot.py:161-190`` reads (``data.json`` = list of segments, each a list of records).

Different by necessity (no pyroomacoustics / CMU ARCTIC / mir_eval offline): mixtures come from
``overiva_b200.synth.convolutive_mixture`` (seeded Laplacian sources, random RIRs, same SINR / SNR mixing rules,
``overiva_sim.py:163-192``) and SDR / SIR from ``overiva_b200.metrics`` (``bss_eval``, one-tap distortion, by default;
``"bss_eval_filter_length": 512`` in the parameters selects ``bss_eval_sources``, the restatement of mir_eval's
512-tap metric, for the initial / final scores).  ``runtime`` is the wall time of the
batched algorithm call divided by the batch size (seconds per mixture, STFT excluded like ``overiva_sim.py:293-320``).
With ``monitor_convergence`` (``overiva_sim.py:272-284``) every mixture runs on its own with the convergence callback
-- scored on the device by ``overiva_b200.monitor`` -- and ``sdr`` / ``sir`` hold one entry per callback plus the final
evaluation; otherwise ``[initial, final]``.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import metrics, synth

# the values of the reference's overiva_sim_config.json that this driver reads (own defaults: shorter runs)
DEFAULT_PARAMETERS = {
    "name": "overiva_b200_sweep",
    "n_repeat": 4,
    "seed": 12345,
    "n_targets_list": [1, 2],
    "n_mics_list": [2, 4, 6],
    "rt60_list": {"0.3": {}},
    "sinr_list": [10],
    "snr": 60,
    "fs": 16000,
    "duration": 5.0,
    "n_interferers": 10,
    "ref_mic": 0,
    "monitor_convergence": False,
    "stft_params": {"framesize": 4096},
    "algorithm_kwargs": {
        "auxiva_laplace": {"algo": "auxiva", "kwargs": {"n_iter": 20, "proj_back": True, "model": "laplace"}},
        "overiva_laplace": {"algo": "overiva",
                            "kwargs": {"n_iter": 20, "proj_back": True, "init_eig": False, "model": "laplace"}},
        "auxiva_pca_laplace": {"algo": "auxiva_pca", "kwargs": {"n_iter": 20, "proj_back": True, "model": "laplace"}},
    },
    "overdet_algos": ["overiva_laplace", "overiva_gauss", "auxiva_pca_laplace", "auxiva_pca_gauss", "ogive_laplace",
                      "ogive_laplace_eig"],
}


def generate_arguments(parameters):
    """``overiva_sim.py:356-392`` without the wav-file sampling: one ``[n_targets, n_mics, rt60, sinr, seed]`` per
    simulated mixture, in the reference's loop order, seeds drawn from ``np.random.seed(parameters['seed'])``."""
    state = np.random.get_state()
    np.random.seed(parameters["seed"])
    np.random.randint(2**32, dtype=np.uint32)  # the reference draws the sample-selection seed first (:363)
    args = []
    for n_targets in parameters["n_targets_list"]:
        for n_mics in parameters["n_mics_list"]:
            if n_targets > n_mics:  # "we don't do underdetermined" (:377-378)
                continue
            for rt60 in parameters["rt60_list"].keys():
                for sinr in parameters["sinr_list"]:
                    for _ in range(parameters["n_repeat"]):
                        seed = int(np.random.randint(2**32, dtype=np.uint32))
                        args.append([n_targets, n_mics, rt60, sinr, seed])
    np.random.set_state(state)
    return args


def algorithms_for(parameters, n_targets):
    """(full_name, algo, kwargs) that ``one_loop`` would run for this case (``overiva_sim.py:245-256,316-317``)."""
    out = []
    for full_name, params in parameters["algorithm_kwargs"].items():
        name = params["algo"]
        if name == "auxiva_pca" and n_targets == 1:
            continue
        if name == "ogive" and n_targets != 1:
            continue
        if name not in ("auxiva", "auxiva_pca", "overiva", "ogive", "ilrma"):
            continue
        out.append((full_name, name, dict(params["kwargs"])))
    return out


def make_mixture(parameters, arg):
    """mix (N, M), ref (n_targets + 1, N, M): target images and the background (``overiva_sim.py:194-197``)."""
    n_targets, n_mics, rt60, sinr, seed = arg
    mix, images = synth.convolutive_mixture(seed % (2**32), n_mics, n_targets, duration=parameters["duration"],
                                            fs=parameters["fs"], n_interferers=parameters["n_interferers"],
                                            sinr_db=float(sinr), snr_db=float(parameters["snr"]), rt60=float(rt60))
    ref = np.concatenate([images, (mix - images.sum(axis=0))[None]], axis=0)
    return mix, ref


def _noise_channel(m, noise_seed):
    return np.random.default_rng(noise_seed).standard_normal(m)  # "fill this to compare to background"


def evaluate(y, ref, n_targets, framesize, reorder, noise_seed=0, flen=1):
    """``convergence_callback`` (``overiva_sim.py:210-232``): y (N', J) time-domain outputs -> (sdr, sir) lists of
    length n_targets, scored at the reference microphone 0.  ``flen``: taps of the allowed distortion filter -- 1 is
    the fast one-tap metric, 512 is what ``mir_eval.separation.bss_eval_sources`` uses (``metrics.bss_eval_sources``)."""
    y = np.asarray(y, dtype=np.float64)
    if reorder:
        y = y[:, np.argsort(np.std(y, axis=0))[::-1]]
    half = framesize // 2
    m = int(min(y.shape[0] - half, ref.shape[1]))
    est = np.zeros((n_targets + 1, m))
    est[:n_targets] = y[half : m + half, :n_targets].T
    est[n_targets] = _noise_channel(m, noise_seed)
    if flen > 1:
        sdr, sir, _, _ = metrics.bss_eval_sources(ref[: n_targets + 1, :m, 0], est, flen=flen)
    else:
        sdr, sir, _ = metrics.bss_eval(ref[: n_targets + 1, :m, 0], est)
    return sdr[:n_targets].tolist(), sir[:n_targets].tolist()


class GpuEngine:
    """Device side of the sweep: one batch of same-shape mixtures per call."""

    def __init__(self, framesize, device=None):
        import torch

        from . import core, stft

        self.torch, self.core, self.stft = torch, core, stft
        self.device = core._require_cuda(device)
        self.L = int(framesize)
        self.hop = self.L // 2
        self.win_a = stft.hann(self.L)
        self.win_s = stft.compute_synthesis_window(self.win_a, self.hop)

    def analysis(self, mixes):
        """mixes (B, N, M) numpy -> device spectra (B, T, F, M); the L - hop zeros in front reproduce the delay the
        reference's evaluation removes (``overiva_sim.py:224-226``)."""
        x = self.torch.from_numpy(np.ascontiguousarray(mixes)).to(self.device)
        return self.stft.analysis(x, self.L, self.hop, win=self.win_a, pad_front=self.L - self.hop)

    def synthesis(self, Y):
        return self.stft.synthesis(Y, self.L, self.hop, win=self.win_s).cpu().numpy()

    def run_monitored(self, algo, Xb, n_targets, kwargs, ref, reorder, noise_seed):
        """One mixture with the reference's convergence callback (``overiva_sim.py:272-284``) kept on the device:
        every 10 epochs (100 for ogive) the estimate is synthesised and scored without leaving the GPU (one Gram
        kernel; ``overiva_b200.monitor``).  -> (Y device (T, F, K), seconds, sdr list, sir list)."""
        torch, core = self.torch, self.core
        from . import monitor

        half = self.L // 2
        refd = torch.from_numpy(np.ascontiguousarray(ref[: n_targets + 1, :, 0])).to(self.device)
        sdrs, sirs = [], []

        def cb(Y):
            y = self.stft.synthesis(Y, self.L, self.hop, win=self.win_s)  # (N', K) on the device
            if reorder:
                y = y[:, torch.argsort(y.std(dim=0), descending=True)]
            m = int(min(y.shape[0] - half, refd.shape[1]))
            est = torch.empty((n_targets + 1, m), dtype=torch.float64, device=self.device)
            est[:n_targets] = y[half : m + half, :n_targets].T
            est[n_targets] = torch.from_numpy(_noise_channel(m, noise_seed)).to(self.device)
            sdr, sir, _ = monitor.bss_eval_device(refd[:, :m], est)
            sdrs.append(sdr[:n_targets].tolist())
            sirs.append(sir[:n_targets].tolist())

        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        if algo == "auxiva":
            Y = core.overiva(Xb, None, callback=cb, **kwargs)
        elif algo == "overiva":
            Y = core.overiva(Xb, n_targets, callback=cb, **kwargs)
        elif algo == "auxiva_pca":
            Y = core.auxiva_pca(Xb, n_src=n_targets, callback=cb, **kwargs)
        elif algo == "ogive":
            Y = core.ogive(Xb, callback=cb, **kwargs)
        elif algo == "ilrma":  # overiva_sim.py:309-311: pra.bss.ilrma(X_mics, callback=cb, **kwargs)
            from .ilrma import ilrma

            Y = ilrma(Xb, callback=cb, **kwargs)
        else:
            raise ValueError(algo)
        torch.cuda.synchronize(self.device)
        dt = time.perf_counter() - t0
        cb(Y)  # "the last evaluation" (overiva_sim.py:320-330)
        return Y, dt, sdrs, sirs

    def run(self, algo, X, n_targets, kwargs):
        """-> (Y device (B, T, F, K), seconds per mixture, failed (B,) bool).  A mixture whose separation fails
        numerically (singular matrix) is flagged, its estimate is NaN, and the others are unaffected -- the reference
        records NaN for the failing task only (``overiva_sim.py:334-350``)."""
        torch, core = self.torch, self.core
        B = X.shape[0]
        failed = np.zeros(B, dtype=bool)
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        if algo in ("auxiva", "overiva"):
            Y, status = core.overiva_batch(X, None if algo == "auxiva" else n_targets, return_status=True, **kwargs)
            failed = (status & 1) != 0
            if failed.any():
                Y[torch.from_numpy(failed).to(Y.device)] = float("nan")
        elif algo in ("auxiva_pca", "ogive", "ilrma"):
            from .ilrma import ilrma

            outs = []
            for b in range(B):
                try:
                    outs.append(core.auxiva_pca(X[b], n_src=n_targets, **kwargs) if algo == "auxiva_pca"
                                else (core.ogive(X[b], **kwargs) if algo == "ogive" else ilrma(X[b], **kwargs)))
                except np.linalg.LinAlgError:
                    failed[b] = True
                    outs.append(None)
            good = next((o for o in outs if o is not None), None)
            if good is None:
                raise np.linalg.LinAlgError("Singular matrix")
            Y = torch.stack([o if o is not None else torch.full_like(good, float("nan")) for o in outs])
        else:
            raise ValueError(algo)
        torch.cuda.synchronize(self.device)
        return Y, (time.perf_counter() - t0) / B, failed


def run(parameters=None, results_dir=None, batch=64, engine=None, progress=None):
    """Run the sweep; returns the list of segments (one list of records per simulated mixture, the structure
    ``rrtools`` writes to ``data.json``).  With ``results_dir`` the three json files of a reference results
    directory are written there."""
    parameters = dict(DEFAULT_PARAMETERS if parameters is None else parameters)
    framesize = parameters["stft_params"]["framesize"]
    args = generate_arguments(parameters)
    if engine is None:
        engine = GpuEngine(framesize)
    segments = [None] * len(args)
    groups = {}
    for i, a in enumerate(args):
        groups.setdefault((a[0], a[1]), []).append(i)
    overdet = set(parameters.get("overdet_algos", []))

    def reorder(algo_name):
        # overiva_sim.py:220 tests the algorithm's BASE name (`name`, passed at :282 and :329) against overdet_algos,
        # which in the shipped configuration lists FULL names ("overiva_laplace", ...): the outputs are therefore
        # always re-ordered by power.  Kept as is -- records must be comparable with the reference's.
        return algo_name not in overdet

    flen = int(parameters.get("bss_eval_filter_length", 1))  # 512 = mir_eval's bss_eval_sources (host, slower)
    rng_state = np.random.get_state()  # (restored on return: ILRMA chunks reseed the global generator)
    for (n_targets, n_mics), idx in groups.items():
        algos = algorithms_for(parameters, n_targets)
        for c0 in range(0, len(idx), batch):
            chunk = idx[c0 : c0 + batch]
            made = [make_mixture(parameters, args[i]) for i in chunk]
            mixes = np.stack([m for m, _ in made])
            refs = [r for _, r in made]
            n_samples = mixes.shape[1]
            X = engine.analysis(mixes)
            # initial SDR / SIR: the microphone signals themselves as the estimate (overiva_sim.py:236-242)
            y0 = engine.synthesis(X[..., :n_targets])
            init = [evaluate(y0[b], refs[b], n_targets, framesize, True, args[i][4], flen) for b, i in enumerate(chunk)]
            recs = [[] for _ in chunk]
            monitored = bool(parameters.get("monitor_convergence", False))
            for full_name, algo, kwargs in algos:
                if algo == "ilrma":
                    # ILRMA draws its initial NMF factors from numpy's global generator.  The reference seeds it per task
                    # (one_loop starts with np.random.seed(seed), overiva_sim.py:105); here the chunk's first seed does,
                    # so that a sweep is reproducible whatever ran before it
                    np.random.seed(args[chunk[0]][4] % (2**32))
                if monitored:
                    # overiva_sim.py:272-284: the callback scores the estimate every 10 epochs; one mixture at a time
                    # (the batched entry point has no callback), the scoring itself stays on the device
                    for b, i in enumerate(chunk):
                        n_t, n_m, rt60, sinr, seed = args[i]
                        try:
                            _, dt, sdrs, sirs = engine.run_monitored(algo, X[b], n_targets, kwargs, refs[b],
                                                                     reorder(algo), seed)
                        except np.linalg.LinAlgError:
                            dt, sdrs, sirs = float("nan"), [[float("nan")]], [[float("nan")]]
                        recs[b].append({
                            "algorithm": full_name, "n_targets": n_t, "n_mics": n_m, "rt60": rt60, "sinr": sinr,
                            "seed": seed, "sdr": sdrs, "sir": sirs, "runtime": dt, "n_samples": int(n_samples),
                        })
                    if progress:
                        progress(full_name, n_targets, n_mics, len(chunk), recs[-1][-1]["runtime"])
                    continue
                nan_score = ([float("nan")], [float("nan")])
                try:
                    Y, per_mix, failed = engine.run(algo, X, n_targets, kwargs)
                    y = engine.synthesis(Y)
                    # the reference logs a failure and records NaN for THAT task only (:333-349)
                    final = [nan_score if failed[b] else
                             evaluate(y[b], refs[b], n_targets, framesize, reorder(algo), args[i][4], flen)
                             for b, i in enumerate(chunk)]
                except np.linalg.LinAlgError:  # every mixture of the chunk failed
                    per_mix, failed = float("nan"), np.ones(len(chunk), dtype=bool)
                    final = [nan_score] * len(chunk)
                for b, i in enumerate(chunk):
                    n_t, n_m, rt60, sinr, seed = args[i]
                    recs[b].append({
                        "algorithm": full_name, "n_targets": n_t, "n_mics": n_m, "rt60": rt60, "sinr": sinr,
                        "seed": seed, "sdr": [init[b][0], final[b][0]], "sir": [init[b][1], final[b][1]],
                        "runtime": float("nan") if failed[b] else per_mix, "n_samples": int(n_samples),
                    })
                if progress:
                    progress(full_name, n_targets, n_mics, len(chunk), per_mix)
            for b, i in enumerate(chunk):
                segments[i] = recs[b]
    np.random.set_state(rng_state)
    if results_dir:
        os.makedirs(results_dir, exist_ok=True)
        with open(os.path.join(results_dir, "parameters.json"), "w") as f:
            json.dump(parameters, f, indent=2)
        with open(os.path.join(results_dir, "arguments.json"), "w") as f:
            json.dump(args, f, indent=0)
        with open(os.path.join(results_dir, "data.json"), "w") as f:
            json.dump(segments, f)
    return segments


def summarise(segments, fs):
    """Mean final SDR / SIR, improvements and real-time factor per (algorithm, n_targets, n_mics): the table
    ``overiva_sim_plot.py:196-243`` builds."""
    acc = {}
    for seg in segments:
        for r in seg:
            key = (r["algorithm"], r["n_targets"], r["n_mics"])
            sdr_i, sdr_f = np.array(r["sdr"][0]), np.array(r["sdr"][-1])
            sir_i, sir_f = np.array(r["sir"][0]), np.array(r["sir"][-1])
            acc.setdefault(key, []).append([r["runtime"] / r["n_samples"] * fs, np.mean(sdr_f), np.mean(sir_f),
                                            np.mean(sdr_f - sdr_i), np.mean(sir_f - sir_i)])
    rows = []
    for key in sorted(acc):
        v = np.nanmean(np.array(acc[key], dtype=float), axis=0)
        rows.append(dict(zip(("algorithm", "n_targets", "n_mics"), key),
                         **dict(zip(("runtime_per_s", "sdr", "sir", "sdr_improvement", "sir_improvement"),
                                    [float(x) for x in v])), n=len(acc[key])))
    return rows


if __name__ == "__main__":  # python -m overiva_b200.sweep [config.json] [results_dir]
    import sys

    params = DEFAULT_PARAMETERS
    if len(sys.argv) > 1:
        with open(sys.argv[1]) as f:
            params = json.load(f)
        params.setdefault("duration", 15.0)
    out_dir = sys.argv[2] if len(sys.argv) > 2 else None
    segs = run(params, out_dir, progress=lambda *a: print("%-22s targets=%d mics=%d batch=%d  %.4f s/mixture" % a))
    for row in summarise(segs, params["fs"]):
        print(json.dumps(row))
