"""Pin the oracle: compare ``oracle.overiva_oracle`` with the UNMODIFIED reference files run under
``oracle.reference_shim`` on seeded synthetic inputs.  Runs only where ``/root/reference`` exists.

    python -m oracle.validate_against_reference

Prints one line per case with the relative Frobenius errors and exits non-zero if any exceeds
``TOL`` (1e-12; observed <= ~1e-13, pure floating-point re-association noise).
"""
from __future__ import annotations

import sys

import numpy as np

from oracle import overiva_oracle as orc
from oracle import reference_shim as ref
from overiva_b200.synth import small_test_mixture

TOL = 1e-12


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def cases():
    # (name, kind, (M, frame), kwargs).  Inputs are small convolutive mixtures (real STFTs), for which the
    # reference is well conditioned; gauss cases are restricted to shapes where a 1e-15 perturbation of
    # X moves the reference's own W by < 1e-12 (verified when the cases were chosen).
    yield "overiva laplace eye M4K2", "overiva", (4, 64), dict(n_src=2, n_iter=20, model="laplace")
    yield "overiva gauss eye M4K2", "overiva", (4, 64), dict(n_src=2, n_iter=20, model="gauss")
    yield "overiva laplace eig M6K2", "overiva", (6, 32), dict(n_src=2, n_iter=20, init_eig=True)
    yield "overiva gauss eig M4K2", "overiva", (4, 32), dict(n_src=2, n_iter=20, model="gauss", init_eig=True)
    yield "overiva laplace eig M8K2", "overiva", (8, 64), dict(n_src=2, n_iter=20, init_eig=True)
    yield "auxiva laplace M3", "overiva", (3, 64), dict(n_iter=20)
    yield "auxiva gauss M6", "overiva", (6, 32), dict(n_iter=20, model="gauss")
    yield "overiva K1 M5", "overiva", (5, 32), dict(n_src=1, n_iter=20)
    yield "overiva K3 M4 no projback", "overiva", (4, 32), dict(n_src=3, n_iter=15, proj_back=False)
    yield "overiva n_iter=0", "overiva", (4, 32), dict(n_src=2, n_iter=0)
    yield "overiva W0", "overiva_W0", (4, 32), dict(n_src=2, n_iter=12)
    yield "overiva complex64", "overiva_c64", (4, 32), dict(n_src=2, n_iter=20)
    yield "auxiva_pca laplace M5K2", "auxiva_pca", (5, 32), dict(n_src=2, n_iter=20, proj_back=True)
    yield "auxiva_pca gauss M4K4", "auxiva_pca", (4, 64), dict(n_src=4, n_iter=10, proj_back=True, model="gauss")
    yield "ogive demix laplace", "ogive", (4, 32), dict(n_iter=60, update="demix")
    yield "ogive mix gauss", "ogive", (4, 32), dict(n_iter=60, update="mix", model="gauss")
    yield "ogive switching eig", "ogive", (3, 32), dict(n_iter=60, update="switching", init_eig=True)
    yield "ogive early stop", "ogive", (3, 32), dict(n_iter=400, tol=5e-2)


# edge shapes of tests/test_edge_gpu.py (T, F, M, K, model): the GPU tests compare with the oracle on exactly these
# inputs, so the oracle is pinned against the reference on them too
EDGE_TOL = 1e-11
EDGE_SHAPES = [(40, 1, 3, 2, "laplace"), (7, 33, 2, 2, "gauss"), (3, 5, 2, 1, "laplace"), (50, 17, 1, 1, "laplace"),
               (64, 9, 16, 16, "laplace"), (300, 9, 16, 4, "gauss"), (90, 40, 9, 3, "laplace"), (60, 64, 7, 7, "gauss")]


def run_edge(T, F, M, K, model):
    from overiva_b200.synth import stft_domain_mixture

    X = stft_domain_mixture(T * 1000 + F * 10 + M, T, F, M, K, n_interferers=min(3, max(0, M - K)), noise_db=-30.0)
    kw = dict(n_src=K, n_iter=5, model=model)
    Yr, Wr = ref.ref_overiva(X, return_filters=True, **kw)
    Yo, Wo = orc.overiva(X, return_filters=True, **kw)
    return {"Y": rel(Yo, Yr), "W": rel(Wo, Wr)}


def run_case(kind, shape, kwargs, seed):
    M, frame = shape
    X = small_test_mixture(seed, M, 2, frame=frame, hop=frame // 2)
    T, F, _ = X.shape
    kw = dict(kwargs)
    if kind == "overiva_W0":
        rng = np.random.default_rng(seed + 1)
        Ks = kw["n_src"]
        W0 = np.zeros((F, M, Ks), dtype=np.complex128)
        W0[:, :Ks, :] = np.eye(Ks)
        W0 += 0.1 * (rng.standard_normal(W0.shape) + 1j * rng.standard_normal(W0.shape))
        kw["W0"] = W0
        kind = "overiva"
    if kind == "overiva_c64":
        X = X.astype(np.complex64)
        kind = "overiva"
    if kind == "overiva":
        Yr, Wr = ref.ref_overiva(X, return_filters=True, **kw)
        Yo, Wo = orc.overiva(X, return_filters=True, **kw)
        return {"Y": rel(Yo, Yr), "W": rel(Wo, Wr)}
    if kind == "auxiva_pca":
        Yr = ref.ref_auxiva_pca(X, **dict(kw))
        Yo = orc.auxiva_pca(X, **dict(kw))
        return {"Y": rel(Yo, Yr)}
    if kind == "ogive":
        Yr, wr = ref.ref_ogive(X, return_filters=True, **kw)
        Yo, wo = orc.ogive(X, return_filters=True, **kw)
        return {"Y": rel(Yo, Yr), "W": rel(wo, wr)}
    raise ValueError(kind)


def main():
    if not ref.available():
        print("reference tree not present; nothing to validate")
        return 0
    worst = 0.0
    for i, (name, kind, shape, kwargs) in enumerate(cases()):
        errs = run_case(kind, shape, kwargs, seed=100 + i)
        tol = 2e-5 if kind == "overiva_c64" else TOL
        ok = all(e <= tol for e in errs.values())
        worst = max(worst, *(e / tol for e in errs.values()))
        print("%-32s %s %s" % (name, " ".join("%s=%.2e" % kv for kv in errs.items()), "ok" if ok else "FAIL"))
    for T, F, M, K, model in EDGE_SHAPES:
        errs = run_edge(T, F, M, K, model)
        # the 16- and 9-channel cases amplify a 1e-15 perturbation of X to ~1e-12 on W in the reference itself
        ok = all(e <= EDGE_TOL for e in errs.values())
        worst = max(worst, *(e / EDGE_TOL for e in errs.values()))
        print("%-32s %s %s" % ("edge T%d F%d M%d K%d %s" % (T, F, M, K, model),
                               " ".join("%s=%.2e" % kv for kv in errs.items()), "ok" if ok else "FAIL"))
    # callback cadence (overiva.py:142: epochs 0, 10, 20, ... before that epoch's update)
    X = small_test_mixture(7, 3, 2, n_samples=1000, frame=16, hop=8)
    got_r, got_o = [], []
    ref.ref_overiva(X, n_src=2, n_iter=25, callback=lambda Y: got_r.append(Y.copy()))
    orc.overiva(X, n_src=2, n_iter=25, callback=lambda Y: got_o.append(Y.copy()))
    cb_ok = len(got_r) == len(got_o) == 3 and all(rel(a, b) <= TOL for a, b in zip(got_o, got_r))
    print("%-32s calls=%d/%d %s" % ("callback cadence", len(got_o), len(got_r), "ok" if cb_ok else "FAIL"))
    if worst > 1.0 or not cb_ok:
        print("ORACLE DOES NOT MATCH THE REFERENCE")
        return 1
    print("oracle == reference within tolerance (worst error / tol = %.3g)" % worst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
