"""CPU tests: the numpy oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py), and -- where /root/reference is mounted -- against the reference itself."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle import overiva_oracle as orc
from oracle import reference_shim


def _run_oracle(case):
    kw = dict(case["kwargs"])
    if "W0" in case:
        kw["W0"] = case["W0"]
    fn = case["fn"]
    if fn == "overiva":
        Y, W = orc.overiva(case["X"], return_filters=True, **kw)
        return Y, W
    if fn == "auxiva_pca":
        return orc.auxiva_pca(case["X"], **kw), None
    if fn == "ogive":
        return orc.ogive(case["X"], return_filters=True, **kw)
    raise ValueError(fn)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    case = load_golden(name)
    Y, W = _run_oracle(case)
    # tolerance: 1e-10 (the north-star fp64 bar) or 1e4 x the reference's own sensitivity to a
    # 1-ulp-level perturbation of X, whichever is larger (only the complex64 case needs the latter)
    tol = max(1e-10, 1e2 * case["sens"])
    assert Y.shape == case["Y"].shape and Y.dtype == case["Y"].dtype
    assert rel_err(Y, case["Y"]) <= tol
    if W is not None:
        assert W.shape == case["W"].shape
        assert rel_err(W, case["W"]) <= tol


def test_projection_back_formula():
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((30, 5, 2)) + 1j * rng.standard_normal((30, 5, 2))
    ref = rng.standard_normal((30, 5)) + 1j * rng.standard_normal((30, 5))
    Y[:, 3, 1] = 0.0  # zero denominator -> z = 1
    z = orc.projection_back(Y, ref)
    for f in range(5):
        for k in range(2):
            den = np.sum(np.abs(Y[:, f, k]) ** 2)
            want = 1.0 if den == 0 else np.sum(np.conj(ref[:, f]) * Y[:, f, k]) / den
            assert abs(z[f, k] - want) < 1e-13
    # least-squares property: conj(z) y is the projection of ref on y
    f, k = 1, 0
    resid = ref[:, f] - np.conj(z[f, k]) * Y[:, f, k]
    assert abs(np.vdot(Y[:, f, k], resid)) < 1e-10


def test_projection_back_via_covariance_identity():
    """z_k = (w^H C e_0) / (w^H C w): the identity the CUDA path uses instead of a pass over Y."""
    case = load_golden("overiva_laplace_m4k2")
    X = case["X"]
    Y, W = orc.overiva(X, return_filters=True, proj_back=False, **{k: v for k, v in case["kwargs"].items() if k != "proj_back"})
    z = orc.projection_back(Y, X[:, :, 0])
    C = orc.input_covariance(X)
    num = np.einsum("fmk,fm->fk", np.conj(W), C[:, :, 0])
    den = np.einsum("fmk,fmn,fnk->fk", np.conj(W), C, W).real
    assert rel_err(num / den, z) < 1e-12


def test_step_functions_compose_to_overiva():
    case = load_golden("overiva_gauss_m4k2")
    X, kw = case["X"], case["kwargs"]
    K = kw["n_src"]
    C = orc.input_covariance(X)
    W_hat = orc.init_demixing(C, K)
    Xf = np.ascontiguousarray(X.swapaxes(0, 1))
    for _ in range(kw["n_iter"]):
        orc.iterate_once(Xf, W_hat, C, K, kw["model"])
    assert rel_err(W_hat[:, :, :K], case["W"]) < 1e-10


def test_frequency_sharded_iteration_equals_full():
    """Splitting bins across shards with a sum of the partial r2 statistics reproduces the full update
    (the decomposition the multi-GPU path relies on: SURVEY.md section 8e)."""
    case = load_golden("overiva_laplace_m6k2")
    X, kw = case["X"], case["kwargs"]
    K, F = kw["n_src"], X.shape[1]
    C = orc.input_covariance(X)
    Xf = np.ascontiguousarray(X.swapaxes(0, 1))
    full = orc.init_demixing(C, K)
    halves = [slice(0, F // 2), slice(F // 2, F)]
    parts = [orc.init_demixing(C[s], K) for s in halves]
    for _ in range(5):
        orc.iterate_once(Xf, full, C, K, "laplace")
        r2 = sum(orc.demix_power(Xf[s], p[:, :, :K]) for s, p in zip(halves, parts))
        for s, p in zip(halves, parts):
            orc.iterate_once(Xf[s], p, C[s], K, "laplace", n_freq_total=F, r2=r2)
    assert rel_err(np.concatenate(parts, axis=0), full) < 1e-12


@pytest.mark.skipif(not reference_shim.available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference():
    from oracle.validate_against_reference import main

    assert main() == 0
