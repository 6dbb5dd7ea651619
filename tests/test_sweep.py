"""Batched sweep driver (reference: overiva_sim.py + rrtools).

CPU: argument enumeration / seeds, algorithm selection rules, record schema and result files, with a numpy stand-in
engine built on the oracle.  GPU: the device engine produces the same records as the stand-in."""
import json
import os

import numpy as np
import pytest

from oracle import ilrma_oracle as ilo
from oracle import overiva_oracle as orc
from oracle import stft_oracle as so
from overiva_b200 import sweep

PARAMS = {
    "name": "test", "n_repeat": 2, "seed": 7, "n_targets_list": [1, 2, 3], "n_mics_list": [2, 3],
    "rt60_list": {"0.02": {}}, "sinr_list": [10], "snr": 60, "fs": 8000, "duration": 0.4, "n_interferers": 3,
    "ref_mic": 0, "monitor_convergence": False, "stft_params": {"framesize": 64},
    "algorithm_kwargs": {
        "auxiva_laplace": {"algo": "auxiva", "kwargs": {"n_iter": 5, "proj_back": True, "model": "laplace"}},
        "overiva_gauss": {"algo": "overiva", "kwargs": {"n_iter": 5, "proj_back": True, "init_eig": False, "model": "gauss"}},
        "auxiva_pca_laplace": {"algo": "auxiva_pca", "kwargs": {"n_iter": 5, "proj_back": True, "model": "laplace"}},
        "ogive_laplace": {"algo": "ogive", "kwargs": {"n_iter": 20, "step_size": 0.1, "tol": 1e-3, "update": "demix",
                                                      "proj_back": True, "model": "laplace", "init_eig": False}},
        "ilrma": {"algo": "ilrma", "kwargs": {"n_iter": 5}},
    },
    "overdet_algos": ["overiva_gauss", "auxiva_pca_laplace", "ogive_laplace"],
}


class OracleEngine:
    """numpy stand-in for sweep.GpuEngine (same interface), built on the oracle."""

    def __init__(self, framesize):
        self.L, self.hop = framesize, framesize // 2
        self.win_a = so.hann(self.L)
        self.win_s = so.compute_synthesis_window(self.win_a, self.hop)

    def analysis(self, mixes):
        return np.stack([so.analysis(m, self.L, self.hop, win=self.win_a, pad_front=self.L - self.hop) for m in mixes])

    def synthesis(self, Y):
        return np.stack([so.synthesis(y, self.L, self.hop, win=self.win_s) for y in Y])

    def run_monitored(self, algo, Xb, n_targets, kwargs, ref, reorder, noise_seed):
        sdrs, sirs = [], []

        def cb(Y):
            sdr, sir = sweep.evaluate(so.synthesis(Y, self.L, self.hop, win=self.win_s), ref, n_targets, self.L, reorder,
                                      noise_seed)
            sdrs.append(sdr)
            sirs.append(sir)

        fn = {"auxiva": lambda x: orc.overiva(x, callback=cb, **kwargs),
              "overiva": lambda x: orc.overiva(x, n_src=n_targets, callback=cb, **kwargs),
              "auxiva_pca": lambda x: orc.auxiva_pca(x, n_src=n_targets, callback=cb, **kwargs),
              "ogive": lambda x: orc.ogive(x, callback=cb, **kwargs),
              "ilrma": lambda x: ilo.ilrma(x, callback=cb, **kwargs)}[algo]
        Y = fn(Xb)
        cb(Y)
        return Y, 0.25, sdrs, sirs

    def run(self, algo, X, n_targets, kwargs):
        fn = {"auxiva": lambda x: orc.overiva(x, **kwargs), "overiva": lambda x: orc.overiva(x, n_src=n_targets, **kwargs),
              "auxiva_pca": lambda x: orc.auxiva_pca(x, n_src=n_targets, **kwargs), "ogive": lambda x: orc.ogive(x, **kwargs),
              "ilrma": lambda x: ilo.ilrma(x, **kwargs)}[algo]
        outs, failed = [], np.zeros(len(X), dtype=bool)
        for b, x in enumerate(X):
            try:
                outs.append(fn(x))
            except np.linalg.LinAlgError:  # only the failing mixture is flagged (overiva_sim.py:334-350)
                failed[b] = True
                outs.append(None)
        good = next(o for o in outs if o is not None)
        return np.stack([o if o is not None else np.full_like(good, np.nan) for o in outs]), 0.5, failed


def test_generate_arguments_follows_the_reference_enumeration():
    args = sweep.generate_arguments(PARAMS)
    # targets-major, then mics; underdetermined (3 targets, 2 mics) skipped; n_repeat entries each
    assert [(a[0], a[1]) for a in args] == [(1, 2)] * 2 + [(1, 3)] * 2 + [(2, 2)] * 2 + [(2, 3)] * 2 + [(3, 3)] * 2
    assert all(a[2] == "0.02" and a[3] == 10 for a in args)
    seeds = [a[4] for a in args]
    assert len(set(seeds)) == len(seeds) and seeds == [a[4] for a in sweep.generate_arguments(PARAMS)]
    # same draws as the reference's loop: np.random.seed(seed); one draw for the file sampling; then one per case
    np.random.seed(7)
    np.random.randint(2**32, dtype=np.uint32)
    assert seeds[0] == int(np.random.randint(2**32, dtype=np.uint32))
    state_before = np.random.get_state()[1][:5].tolist()
    sweep.generate_arguments(PARAMS)
    assert np.random.get_state()[1][:5].tolist() == state_before  # caller's RNG state restored (overiva_sim.py:390)


def test_algorithm_selection_rules():
    names = lambda n: [a[0] for a in sweep.algorithms_for(PARAMS, n)]
    assert names(1) == ["auxiva_laplace", "overiva_gauss", "ogive_laplace", "ilrma"]  # no PCA for a single target
    assert names(2) == ["auxiva_laplace", "overiva_gauss", "auxiva_pca_laplace", "ilrma"]  # OGIVE only for one target
    unknown = dict(PARAMS, algorithm_kwargs={"x": {"algo": "fastica", "kwargs": {}}})
    assert sweep.algorithms_for(unknown, 2) == []  # one_loop's `else: continue` (overiva_sim.py:316-317)


def test_sweep_records_and_files(tmp_path):
    p = dict(PARAMS, n_targets_list=[1, 2], n_repeat=2)
    seen = []
    segs = sweep.run(p, str(tmp_path), batch=3, engine=OracleEngine(64), progress=lambda *a: seen.append(a))
    args = sweep.generate_arguments(p)
    assert len(segs) == len(args) == 8
    for seg, a in zip(segs, args):
        assert [r["algorithm"] for r in seg] == [x[0] for x in sweep.algorithms_for(p, a[0])]
        for r in seg:
            assert set(r) == {"algorithm", "n_targets", "n_mics", "rt60", "sinr", "seed", "sdr", "sir", "runtime",
                              "n_samples"}
            assert (r["n_targets"], r["n_mics"], r["rt60"], r["sinr"], r["seed"]) == tuple(a)
            assert len(r["sdr"]) == 2 and len(r["sdr"][0]) == len(r["sdr"][1]) == a[0]
            assert r["n_samples"] == 3200 and r["runtime"] == 0.5
            assert np.all(np.isfinite(r["sdr"])) and np.all(np.isfinite(r["sir"]))
    # initial values are shared by every algorithm of a mixture (overiva_sim.py:285-287)
    assert segs[0][0]["sdr"][0] == segs[0][1]["sdr"][0]
    # files in the layout overiva_sim_plot.py:161-190 reads
    data = json.load(open(os.path.join(tmp_path, "data.json")))
    records = []
    for seg in data:
        records += seg
    assert len(records) == sum(len(s) for s in segs) and records[0] == segs[0][0]
    assert json.load(open(os.path.join(tmp_path, "parameters.json")))["fs"] == 8000
    assert json.load(open(os.path.join(tmp_path, "arguments.json"))) == args
    rows = sweep.summarise(segs, p["fs"])
    assert {r["algorithm"] for r in rows} == {"auxiva_laplace", "overiva_gauss", "auxiva_pca_laplace", "ogive_laplace", "ilrma"}
    assert all(r["runtime_per_s"] == pytest.approx(0.5 / 3200 * 8000) for r in rows)
    assert seen and seen[0][0] == "auxiva_laplace"


def test_separation_improves_sir():
    p = dict(PARAMS, n_targets_list=[2], n_mics_list=[3], n_repeat=3, duration=1.0,
             algorithm_kwargs={"overiva_laplace": {"algo": "overiva", "kwargs": {"n_iter": 30, "proj_back": True,
                                                                                "model": "laplace"}}},
             overdet_algos=["overiva_laplace"])
    rows = sweep.summarise(sweep.run(p, engine=OracleEngine(64)), p["fs"])
    assert rows[0]["sir_improvement"] > 3.0


def test_filter_based_metric_option():
    """bss_eval_filter_length > 1 scores with the mir_eval-style filter metric: never lower SDR than the one-tap metric
    (the allowed distortion is a superset), same record layout."""
    algs = {"overiva_laplace": {"algo": "overiva", "kwargs": {"n_iter": 10, "proj_back": True, "model": "laplace"}}}
    p1 = dict(PARAMS, n_targets_list=[2], n_mics_list=[3], n_repeat=1, algorithm_kwargs=algs, overdet_algos=["overiva_laplace"])
    p2 = dict(p1, bss_eval_filter_length=16)
    r1 = sweep.run(p1, engine=OracleEngine(64))[0][0]
    r2 = sweep.run(p2, engine=OracleEngine(64))[0][0]
    assert len(r2["sdr"]) == 2 and len(r2["sdr"][1]) == 2
    assert np.all(np.array(r2["sdr"][1]) >= np.array(r1["sdr"][1]) - 1e-9)


def test_monitor_convergence_mode():
    """overiva_sim.py:272-284: with monitor_convergence the sdr / sir lists hold one entry per callback (epochs 0, 10,
    ...) plus the final evaluation, instead of [initial, final]."""
    algs = {"overiva_laplace": {"algo": "overiva", "kwargs": {"n_iter": 25, "proj_back": True, "model": "laplace"}},
            "auxiva_laplace": {"algo": "auxiva", "kwargs": {"n_iter": 12, "proj_back": True, "model": "laplace"}}}
    p = dict(PARAMS, n_targets_list=[2], n_mics_list=[3], n_repeat=2, monitor_convergence=True, algorithm_kwargs=algs,
             overdet_algos=["overiva_laplace"], duration=1.0)
    segs = sweep.run(p, engine=OracleEngine(64))
    assert len(segs) == 2
    for seg in segs:
        by = {r["algorithm"]: r for r in seg}
        assert len(by["overiva_laplace"]["sdr"]) == 3 + 1 and len(by["auxiva_laplace"]["sdr"]) == 2 + 1
        assert all(len(v) == 2 for v in by["overiva_laplace"]["sir"])
        assert by["overiva_laplace"]["runtime"] == 0.25
    # over the sweep the separation improves between the first callback (before any update) and the final evaluation
    assert np.mean([np.mean(s[0]["sir"][-1]) - np.mean(s[0]["sir"][0]) for s in segs]) > 0
    rows = sweep.summarise(segs, p["fs"])  # first entry = before the first update, last = final: improvements defined
    assert all(np.isfinite(r["sir_improvement"]) for r in rows)


@pytest.mark.gpu
def test_gpu_monitored_sweep_matches_the_oracle_engine():
    algs = {"overiva_laplace": {"algo": "overiva", "kwargs": {"n_iter": 12, "proj_back": True, "model": "laplace"}}}
    p = dict(PARAMS, n_targets_list=[2], n_mics_list=[3], n_repeat=2, monitor_convergence=True, algorithm_kwargs=algs,
             overdet_algos=["overiva_laplace"])
    ref = sweep.run(p, engine=OracleEngine(64))
    got = sweep.run(p)
    for sg, sr in zip(got, ref):
        for rg, rr in zip(sg, sr):
            assert len(rg["sdr"]) == len(rr["sdr"]) == 3
            assert np.allclose(rg["sdr"], rr["sdr"], atol=1e-6) and np.allclose(rg["sir"], rr["sir"], atol=1e-6)


@pytest.mark.gpu
def test_gpu_engine_matches_the_oracle_engine():
    p = dict(PARAMS, n_targets_list=[1, 2], n_mics_list=[3], n_repeat=2)
    ref = sweep.run(p, engine=OracleEngine(64))
    got = sweep.run(p, batch=2)
    assert len(got) == len(ref)
    for sg, sr in zip(got, ref):
        for rg, rr in zip(sg, sr):
            assert rg["algorithm"] == rr["algorithm"] and rg["seed"] == rr["seed"]
            assert np.allclose(rg["sdr"], rr["sdr"], atol=1e-6) and np.allclose(rg["sir"], rr["sir"], atol=1e-6)
            assert rg["runtime"] > 0


@pytest.mark.gpu
def test_one_failing_mixture_gets_nan_and_the_rest_of_the_chunk_is_scored(monkeypatch):
    """overiva_sim.py:334-350: a task that raises is recorded with NaN and the farm carries on.  One rank-deficient
    mixture (a duplicated microphone) inside a GPU chunk must not take the other mixtures of the chunk with it."""
    algs = {"overiva_laplace": {"algo": "overiva", "kwargs": {"n_iter": 8, "proj_back": True, "model": "laplace"}},
            "auxiva_laplace": {"algo": "auxiva", "kwargs": {"n_iter": 8, "proj_back": True, "model": "laplace"}}}
    p = dict(PARAMS, n_targets_list=[2], n_mics_list=[3], n_repeat=4, algorithm_kwargs=algs, overdet_algos=[])
    real = sweep.make_mixture
    calls = []

    def broken(parameters, arg):
        mix, ref = real(parameters, arg)
        calls.append(arg)
        if len(calls) % 4 == 2:  # the second mixture of every run of four
            mix = mix.copy()
            mix[:, 2] = mix[:, 0]
        return mix, ref

    monkeypatch.setattr(sweep, "make_mixture", broken)
    got = sweep.run(p, batch=4)
    calls.clear()
    ref = sweep.run(p, engine=OracleEngine(64))
    assert len(got) == len(ref) == 4
    for i, (sg, sr) in enumerate(zip(got, ref)):
        for rg, rr in zip(sg, sr):
            failed = i == 1
            assert np.isnan(rg["runtime"]) == failed
            assert bool(np.all(np.isnan(rg["sdr"][1]))) == failed and bool(np.all(np.isnan(rr["sdr"][1]))) == failed
            if not failed:
                assert np.allclose(rg["sdr"], rr["sdr"], atol=1e-6) and np.allclose(rg["sir"], rr["sir"], atol=1e-6)
