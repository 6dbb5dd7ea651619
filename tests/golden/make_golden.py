"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

    python tests/golden/make_golden.py          # needs /root/reference (build container only)
    python tests/golden/make_golden.py --only=name1,name2     # (re)generate the named cases only

Each ``<case>.npz`` holds a small seeded input ``X`` (a real STFT of a synthetic convolutive mixture,
``overiva_b200.synth.small_test_mixture``), the keyword arguments of the call (JSON) and the outputs
of the reference function run under ``oracle.reference_shim`` (``Y`` and, where the function returns
them, the filters ``W``).  The reference tree does not exist on the GPU box, so the fixtures -- not the
reference -- are what ``tests/`` read at run time.

A case is only written if the reference itself is well conditioned on it: a 1e-15 relative perturbation
of X must move the reference's W and Y by < 1e-12 (otherwise the next seed is tried); the measured
sensitivity is stored in the fixture (``sens``) so tests can state their tolerance against it.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_shim as ref  # noqa: E402
from overiva_b200.synth import small_test_mixture, stft_domain_mixture  # noqa: E402

# name, function, M, (n_samples, frame), dtype, kwargs
CASES = [
    ("overiva_laplace_m4k2", "overiva", 4, (1000, 32), "c128", dict(n_src=2, n_iter=20, model="laplace")),
    ("overiva_gauss_m4k2", "overiva", 4, (1000, 32), "c128", dict(n_src=2, n_iter=20, model="gauss")),
    ("overiva_laplace_eig_m6k2", "overiva", 6, (1000, 32), "c128", dict(n_src=2, n_iter=20, init_eig=True)),
    ("overiva_laplace_m6k2", "overiva", 6, (1000, 32), "c128", dict(n_src=2, n_iter=20)),
    ("overiva_gauss_eig_m8k2", "overiva", 8, (2400, 32), "c128", dict(n_src=2, n_iter=20, model="gauss", init_eig=True)),
    ("auxiva_laplace_m3", "overiva", 3, (1000, 32), "c128", dict(n_iter=20)),
    ("auxiva_laplace_m6", "overiva", 6, (1000, 32), "c128", dict(n_iter=20)),
    ("auxiva_gauss_m4", "overiva", 4, (1000, 32), "c128", dict(n_iter=20, model="gauss")),
    ("overiva_k1_m5", "overiva", 5, (1000, 32), "c128", dict(n_src=1, n_iter=20)),
    ("overiva_k3_m4_noprojback", "overiva", 4, (1000, 32), "c128", dict(n_src=3, n_iter=15, proj_back=False)),
    ("overiva_w0_m4k2", "overiva_W0", 4, (1000, 32), "c128", dict(n_src=2, n_iter=12)),
    ("overiva_c64_m4k2", "overiva", 4, (1000, 32), "c64", dict(n_src=2, n_iter=20)),
    ("overiva_niter0_m4k2", "overiva", 4, (600, 32), "c128", dict(n_src=2, n_iter=0)),
    ("overiva_m2", "overiva", 2, (1000, 16), "c128", dict(n_iter=20)),
    ("auxiva_pca_laplace_m5k2", "auxiva_pca", 5, (1000, 32), "c128", dict(n_src=2, n_iter=20, proj_back=True)),
    ("auxiva_pca_gauss_m4k4", "auxiva_pca", 4, (1000, 32), "c128", dict(n_src=4, n_iter=10, proj_back=True, model="gauss")),
    ("ogive_demix_laplace_m4", "ogive", 4, (1000, 32), "c128", dict(n_iter=60, update="demix")),
    ("ogive_mix_gauss_m4", "ogive", 4, (1000, 32), "c128", dict(n_iter=60, update="mix", model="gauss")),
    ("ogive_switching_eig_m3", "ogive", 3, (1000, 32), "c128", dict(n_iter=60, update="switching", init_eig=True)),
    ("ogive_earlystop_m3", "ogive", 3, (1000, 32), "c128", dict(n_iter=400, tol=5e-2)),
    # many-channel shapes (added with `--only`: the fixtures above are not regenerated).  ("stft", T, F) draws X in
    # the STFT domain (overiva_b200.synth.stft_domain_mixture) -- a 2-target convolutive mixture seen by 8-16
    # microphones is nearly rank deficient; these cases accept a sensitivity of 1e-11
    ("auxiva_laplace_m8", "overiva", 8, ("stft", 120, 33), "c128", dict(n_iter=10)),
    ("overiva_laplace_m8k4", "overiva", 8, ("stft", 120, 33), "c128", dict(n_src=4, n_iter=20)),
    ("overiva_laplace_m16k4", "overiva", 16, ("stft", 300, 17), "c128", dict(n_src=4, n_iter=6)),
    ("auxiva_laplace_m16", "overiva", 16, ("stft", 64, 9), "c128", dict(n_iter=5)),
    ("overiva_laplace_m9k3", "overiva", 9, ("stft", 90, 40), "c128", dict(n_src=3, n_iter=10)),
]


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def call(fn, X, kw, W0):
    kw = dict(kw)
    if W0 is not None:
        kw["W0"] = W0
    if fn in ("overiva", "overiva_W0"):
        Y, W = ref.ref_overiva(X, return_filters=True, **kw)
        return Y, np.ascontiguousarray(W)
    if fn == "auxiva_pca":
        return ref.ref_auxiva_pca(X, **kw), None
    if fn == "ogive":
        Y, W = ref.ref_ogive(X, return_filters=True, **kw)
        return Y, np.ascontiguousarray(W)
    raise ValueError(fn)


def main():
    if not ref.available():
        raise SystemExit("reference tree not present: fixtures can only be generated in the build container")
    prng = np.random.default_rng(2024)
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):
            only = set(a[len("--only="):].split(","))
    for idx, (name, fn, M, shape, dt, kw) in enumerate(CASES):
        if only is not None and name not in only:
            continue
        if only is not None:
            prng = np.random.default_rng(2024 + idx)
        sens_max = 1e-11 if shape[0] == "stft" else 1e-12
        for attempt in range(20):
            seed = 1000 + 37 * idx + attempt
            if shape[0] == "stft":
                K = kw.get("n_src") or M
                X = stft_domain_mixture(seed, shape[1], shape[2], M, K, n_interferers=min(3, max(0, M - K)),
                                        noise_db=-30.0)
            else:
                n_samples, frame = shape
                X = small_test_mixture(seed, M, 2, n_samples=n_samples, frame=frame, hop=frame // 2)
            if dt == "c64":
                X = X.astype(np.complex64)
            W0 = None
            if fn == "overiva_W0":
                r = np.random.default_rng(seed + 1)
                K = kw["n_src"]
                W0 = np.zeros((X.shape[1], M, K), dtype=np.complex128)
                W0[:, :K, :] = np.eye(K)
                W0 += 0.1 * (r.standard_normal(W0.shape) + 1j * r.standard_normal(W0.shape))
            Y, W = call(fn, X, kw, W0)
            eps = 1e-15 if dt == "c128" else 1e-7
            Xp = (X * (1 + eps * prng.standard_normal(X.shape))).astype(X.dtype)
            Yp, Wp = call(fn, Xp, kw, W0)
            sens = max(rel(Yp, Y), rel(Wp, W) if W is not None else 0.0)
            if np.all(np.isfinite(Y)) and sens < sens_max * (eps / 1e-15):
                break
        else:
            raise SystemExit("no well-conditioned seed found for " + name)
        out = dict(X=X, Y=Y, kwargs=json.dumps(kw), fn=fn.replace("_W0", ""), seed=seed, sens=sens)
        if W is not None:
            out["W"] = W
        if W0 is not None:
            out["W0"] = W0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-28s X%s seed=%d sens=%.1e" % (name, X.shape, seed, sens))


if __name__ == "__main__":
    main()
