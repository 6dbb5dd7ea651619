"""GPU parity tests, kernel by kernel, through the C ABI: every CUDA step against the numpy oracle's
restatement of the reference step it replaces (oracle/overiva_oracle.py).  fp64 bar: 1e-12 here (single
steps; the end-to-end bar of 1e-10 after 20 iterations is in test_api_gpu.py)."""
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import overiva_oracle as orc
from overiva_b200.synth import small_test_mixture

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if torch.cuda.is_available():
    import gpu_util as G
    from overiva_b200 import _lib as L


def _mix(seed, M, T_samples=1500, frame=64, dtype=np.complex128, B=1):
    Xs = [small_test_mixture(seed + b, M, 2, n_samples=T_samples, frame=frame, hop=frame // 2) for b in range(B)]
    return np.stack(Xs).astype(dtype)


# shapes cover: F below / at / above a multiple of 32 bins, odd channel counts, float storage
@pytest.mark.parametrize("M,n_samples,dtype", [(4, 1500, np.complex128), (3, 1000, np.complex64), (6, 4200, np.complex128),
                                                (16, 2300, np.complex128), (5, 4900, np.complex64), (1, 700, np.complex128)])
def test_relayout(M, n_samples, dtype):
    X = _mix(1, M, n_samples, 64 if M % 2 == 0 else 32, dtype, B=2)  # F = 33 (two groups, one ragged) or 17
    got = G.grouped(X).cpu().numpy().view(dtype)
    want = G.grouped_expected(X)
    assert got.shape == want.shape
    assert np.array_equal(got, want)  # pure data movement: bit exact


@pytest.mark.parametrize("M,K,n_samples,dtype", [
    (4, 2, 1500, np.complex128), (6, 2, 2000, np.complex128), (6, 6, 1200, np.complex128), (2, 1, 900, np.complex128),
    (3, 3, 4200, np.complex128), (8, 2, 5000, np.complex128), (16, 4, 2300, np.complex128), (5, 5, 1500, np.complex128),
    (7, 3, 1500, np.complex64), (4, 2, 4500, np.complex64), (12, 8, 1000, np.complex128), (9, 7, 2100, np.complex128),
])
def test_weighted_covariance(M, K, n_samples, dtype):
    B = 2
    X = _mix(2, M, n_samples, 64 if M % 2 == 0 else 32, dtype, B=B)
    _, T, F, _ = X.shape
    rng = np.random.default_rng(5)
    phi = rng.gamma(1.0, 1.0, size=(B, K, T)) + 0.01
    code = G.code_of(dtype)
    Xg = G.grouped(X)
    V = G.weighted_cov(Xg, phi, B, T, F, M, K, code)
    X128 = X.astype(np.complex128)
    tol = 1e-12 if dtype == np.complex128 else 1e-12  # products are formed in fp64 in both modes
    for b in range(B):
        Xf = np.ascontiguousarray(X128[b].swapaxes(0, 1))
        for k in range(K):
            want = orc.weighted_covariance(Xf, phi[b, k])
            assert rel_err(V[b, :, k], want) < tol, (b, k)
    # Hermitian by construction, real diagonal
    assert np.array_equal(V, np.conj(V.swapaxes(-1, -2)))
    # plain covariance (phi == NULL)
    C = G.weighted_cov(Xg, None, B, T, F, M, 1, code)
    for b in range(B):
        assert rel_err(C[b, :, 0], orc.input_covariance(X128[b])) < tol


def test_weighted_covariance_split_rows():
    """few bins, many frames: the frames of a group are split over teams and combined atomically"""
    M, K = 4, 2
    X = _mix(3, M, 40000, 16, np.complex128, B=1)[:, :, :3]  # F = 3 bins, T ~ 5000 frames
    _, T, F, _ = X.shape
    rng = np.random.default_rng(6)
    phi = rng.gamma(1.0, 1.0, size=(1, K, T)) + 0.01
    Xg = G.grouped(X)
    V = G.weighted_cov(Xg, phi, 1, T, F, M, K, L.C128)
    Xf = np.ascontiguousarray(X[0].swapaxes(0, 1))
    for k in range(K):
        assert rel_err(V[0, :, k], orc.weighted_covariance(Xf, phi[0, k])) < 1e-12


@pytest.mark.parametrize("M,n_samples,frame,dtype,B", [(6, 4200, 64, np.complex128, 2), (4, 1500, 64, np.complex128, 1),
                                                       (3, 1000, 32, np.complex128, 2), (1, 700, 64, np.complex128, 1),
                                                       (8, 30000, 64, np.complex128, 1), (6, 2500, 64, np.complex64, 2),
                                                       (5, 2500, 62, np.complex64, 1), (7, 900, 16, np.complex128, 3)])
def test_relayout_cov_fused_equals_two_kernels(M, n_samples, frame, dtype, B):
    """One pass (relayout + input covariance) == oiva_relayout followed by the unweighted oiva_weighted_cov_ws:
    identical grouped samples, identical covariance (same arithmetic in the same order)."""
    lib = L.load()
    X = _mix(31, M, n_samples, frame, dtype, B=B)
    B, T, F, M = X.shape
    code = G.code_of(X.dtype)
    if not lib.oiva_relayout_cov_supported(F, M, code):
        assert code == L.C64 and (F * M) % 2 == 1
        pytest.skip("complex64 rows of odd length are not 16-byte aligned: the plan keeps the two-kernel path")
    Xd = G.to_dev(X)
    nx = lib.oiva_grouped_bytes(B, T, F, M, code)
    nc = lib.oiva_grouped_cov_bytes(B, F, M, 1)
    ws_bytes = lib.oiva_weighted_cov_scratch_bytes(B, T, F, M, 1)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=G.dev())
    Xg1 = torch.full((nx,), 0xFF, dtype=torch.uint8, device=G.dev())
    Cg1 = torch.full((nc,), 0xFF, dtype=torch.uint8, device=G.dev())
    L.check(lib.oiva_relayout_cov(G.P(Xd), G.P(Xg1), G.P(Cg1), G.P(ws), ws_bytes, B, T, F, M, code, G.stream()),
            "oiva_relayout_cov")
    torch.cuda.synchronize()
    Xg0 = G.grouped(X)
    Cg0 = torch.full((nc,), 0xFF, dtype=torch.uint8, device=G.dev())
    L.check(lib.oiva_weighted_cov_ws(G.P(Xg0), None, G.P(Cg0), G.P(ws), ws_bytes, B, T, F, M, 1, code, G.stream()),
            "oiva_weighted_cov_ws")
    torch.cuda.synchronize()
    assert torch.equal(Xg1, Xg0)
    a, b = Cg1.view(torch.float64).cpu().numpy(), Cg0.view(torch.float64).cpu().numpy()
    assert np.all(np.isfinite(a)) and rel_err(a, b) < 1e-14
    # and against numpy
    C = torch.empty((B, F, 1, M, M), dtype=torch.complex128, device=G.dev())
    L.check(lib.oiva_unpack_cov(G.P(Cg1), G.P(C), B, F, M, 1, G.stream()), "oiva_unpack_cov")
    torch.cuda.synchronize()
    want = np.stack([orc.input_covariance(X[b].astype(np.complex128)) for b in range(B)])
    assert rel_err(C.cpu().numpy()[:, :, 0], want) < (1e-6 if dtype == np.complex64 else 1e-13)
    # no scratch: no frame splitting, same result to rounding
    Cg2 = torch.full((nc,), 0xFF, dtype=torch.uint8, device=G.dev())
    L.check(lib.oiva_relayout_cov(G.P(Xd), G.P(Xg1), G.P(Cg2), None, 0, B, T, F, M, code, G.stream()), "oiva_relayout_cov")
    torch.cuda.synchronize()
    assert rel_err(Cg2.view(torch.float64).cpu().numpy(), b) < 1e-13


def test_weighted_covariance_deterministic_split():
    """With scratch, frame-split partial sums are combined in a fixed order: bit-identical from run to run and equal
    (to rounding) to the atomic combination."""
    lib = L.load()
    X = _mix(21, 5, 30000, 64)
    B, T, F, M = X.shape
    K = 2
    code = G.code_of(X.dtype)
    Xg = G.grouped(X)
    Tp = lib.oiva_frame_pitch(T)
    rng = np.random.default_rng(1)
    ph = np.zeros((B, K, Tp))
    ph[:, :, :T] = rng.gamma(1.0, 1.0, size=(B, K, T)) + 0.05
    phid = G.to_dev(ph)
    nbytes = lib.oiva_grouped_cov_bytes(B, F, M, K)
    ws_bytes = lib.oiva_weighted_cov_scratch_bytes(B, T, F, M, K)
    assert ws_bytes >= 2 * nbytes
    outs = []
    for _ in range(3):
        ws = torch.full((ws_bytes,), 0xFF, dtype=torch.uint8, device=G.dev())
        Vg = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device=G.dev())
        L.check(lib.oiva_weighted_cov_ws(G.P(Xg), G.P(phid), G.P(Vg), G.P(ws), ws_bytes, B, T, F, M, K, code, G.stream()),
                "oiva_weighted_cov_ws")
        torch.cuda.synchronize()
        outs.append(Vg.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    Va = torch.empty((nbytes,), dtype=torch.uint8, device=G.dev())
    L.check(lib.oiva_weighted_cov(G.P(Xg), G.P(phid), G.P(Va), B, T, F, M, K, code, G.stream()), "oiva_weighted_cov")
    torch.cuda.synchronize()
    a, b = outs[0].view(torch.float64).cpu().numpy(), Va.view(torch.float64).cpu().numpy()
    assert rel_err(a, b) < 1e-14
    # a scratch too small for two slots falls back to the atomic path; no scratch needed for large batches
    assert lib.oiva_weighted_cov_scratch_bytes(512, 116, 2049, 6, 2) == 0


@pytest.mark.parametrize("M,K,n_samples,dtype", [(4, 2, 1500, np.complex128), (6, 6, 1200, np.complex128),
                                                  (8, 3, 5000, np.complex128), (16, 4, 2300, np.complex128),
                                                  (5, 1, 1500, np.complex64), (10, 10, 900, np.complex128),
                                                  (7, 7, 2000, np.complex64), (3, 3, 700, np.complex128)])
def test_demix_power(M, K, n_samples, dtype):
    B = 2
    X = _mix(4, M, n_samples, 64 if M % 2 == 0 else 32, dtype, B=B)
    _, T, F, _ = X.shape
    rng = np.random.default_rng(7)
    W = rng.standard_normal((B, F, M, M)) + 1j * rng.standard_normal((B, F, M, M))
    Xg = G.grouped(X)
    r2, part = G.demix_power(Xg, W, B, T, F, M, K, G.code_of(dtype))
    assert np.all(np.isfinite(part))  # every (chunk, k, t) slot was written, padding included
    X128 = X.astype(np.complex128)
    for b in range(B):
        Xf = np.ascontiguousarray(X128[b].swapaxes(0, 1))
        want = orc.demix_power(Xf, W[b][:, :, :K])  # (T, K)
        assert rel_err(r2[b].T, want) < 1e-12


@pytest.mark.parametrize("T", [173, 2049, 5000])  # > 2048 frames: the two-kernel path for long mixtures
@pytest.mark.parametrize("model", ["laplace", "gauss"])
def test_source_model(model, T):
    rng = np.random.default_rng(8)
    B, K, M, F = 3, 2, 4, 65
    r2 = rng.gamma(1.0, 1.0, size=(B, K, T))
    r2[0, 0, 5] = 0.0  # exercises the 1e-15 clamp
    code = {"laplace": L.MODEL_LAPLACE, "gauss": L.MODEL_GAUSS}[model]
    phi, ws = G.source_model(r2, T, F, code)
    for b in range(B):
        r_inv, w_scale = orc.source_model(r2[b].T, model, F)
        assert rel_err(phi[b].T, r_inv) < 1e-14
        assert rel_err(ws[b], 1.0 / w_scale) < 1e-14


@pytest.mark.parametrize("grouped_c", [True, False])  # thread-per-bin sweep (registers: M <= 6, M = 7, 8 with K <= 4; shared memory: other M >= 7) / lane-group-per-bin sweep
@pytest.mark.parametrize("M,K", [(4, 2), (6, 2), (6, 6), (3, 3), (2, 1), (8, 2), (5, 4), (16, 4), (16, 16), (1, 1), (9, 3),
                                 (6, 1), (6, 4), (5, 5), (4, 3), (7, 1), (7, 3), (7, 4), (8, 1), (8, 4), (8, 5), (7, 7),
                                 (8, 8), (12, 12), (13, 2), (16, 15), (10, 1)])
def test_ip_update_sweep(M, K, grouped_c):
    X = _mix(9, M, 1500, 64 if M % 2 == 0 else 32)[0]
    T, F, _ = X.shape
    Cx = orc.input_covariance(X)
    rng = np.random.default_rng(10)
    What = orc.init_demixing(Cx, K)
    # perturb W so the sweep starts from a generic point, then restore the structure (J, -I)
    What[:, :, :K] += 0.2 * (rng.standard_normal((F, M, K)) + 1j * rng.standard_normal((F, M, K)))
    if K < M:
        orc.background_update(What, Cx, K)
    Xf = np.ascontiguousarray(X.swapaxes(0, 1))
    r_inv = rng.gamma(1.0, 1.0, size=(T, K)) + 0.05
    wscale = rng.uniform(0.5, 2.0, size=(1, K))
    V = np.stack([orc.weighted_covariance(Xf, r_inv[:, s]) for s in range(K)], axis=1)  # (F, K, M, M)
    got, status = G.ip_update(What[None], V[None], Cx[None], wscale, K, grouped_c)
    want = What.copy()
    want[:, :, :K] *= wscale[0][None, None, :]
    for s in range(K):
        orc.ip_update_source(want, V[:, s], Cx, s, K)
    assert status == 0
    assert rel_err(got[0], want) < 1e-11


def test_ip_update_flags_singular():
    M, K, F = 4, 2, 5
    rng = np.random.default_rng(11)
    What = np.zeros((1, F, M, M), dtype=np.complex128)
    V = np.zeros((1, F, K, M, M), dtype=np.complex128)  # singular on purpose
    Cx = np.tile(np.eye(M, dtype=np.complex128), (1, F, 1, 1))
    What[0, :, :K, :K] = np.eye(K)
    _, status = G.ip_update(What, V, Cx, None, K)
    assert status & L.STATUS_SINGULAR


@pytest.mark.parametrize("M,K", [(4, 2), (6, 6), (9, 3)])
def test_ip_update_status_is_per_mixture(M, K):
    """Three mixtures, the middle one with zero covariances: only ITS status word is set (overiva_sim.py:334-350
    records NaN for the failing mixture only)."""
    F, B = 37, 3
    rng = np.random.default_rng(12)
    What = np.zeros((B, F, M, M), dtype=np.complex128)
    What[:, :, np.arange(M), np.arange(M)] = np.where(np.arange(M) < K, 1.0, -1.0)
    A = rng.standard_normal((B, F, K, M, 2 * M)) + 1j * rng.standard_normal((B, F, K, M, 2 * M))
    V = A @ np.conj(np.swapaxes(A, -1, -2)) / (2 * M)
    V[1] = 0.0
    Cx = np.tile(np.eye(M, dtype=np.complex128), (B, F, 1, 1))
    for grouped_c in (True, False):
        _, status = G.ip_update(What, V, Cx, None, K, grouped_c)
        assert status[0] == 0 and status[2] == 0, status
        assert status[1] & (L.STATUS_SINGULAR | L.STATUS_NONFINITE), status


@pytest.mark.parametrize("M,K", [(4, 2), (6, 2), (3, 3), (8, 1), (16, 4)])
def test_init_demix_eye_and_w0(M, K):
    X = _mix(12, M, 1200, 32)[0]
    F = X.shape[1]
    Cx = orc.input_covariance(X)
    got, status = G.init_demix(Cx, K, L.INIT_EYE)
    assert status == 0
    assert rel_err(got, orc.init_demixing(Cx, K)) < 1e-12
    rng = np.random.default_rng(13)
    W0 = rng.standard_normal((F, M, K)) + 1j * rng.standard_normal((F, M, K))
    got, status = G.init_demix(Cx, K, L.INIT_W0, W0=W0)
    assert rel_err(got, orc.init_demixing(Cx, K, W0=W0)) < 1e-11


@pytest.mark.parametrize("M,K", [(4, 2), (6, 2), (3, 3), (8, 1), (6, 6), (7, 4), (1, 1), (16, 4), (8, 8)])
def test_init_demix_grouped(M, K):
    """The thread-per-bin initialisation that writes the grouped W_hat directly (identity and W0) against the oracle and
    against oiva_init_demix; shapes outside the thread-per-bin set answer OIVA_ERR_UNSUPPORTED."""
    B = 2
    X = _mix(18, M, 1300, 80, B=B)  # F = 41: a ragged second group
    F = X.shape[2]
    Cx = np.stack([orc.input_covariance(X[b]) for b in range(B)])
    got, status, rc = G.init_demix_grouped(Cx, K)
    if M > 8 or (M > 6 and K > 4):
        assert rc == L.ERR_UNSUPPORTED
        return
    assert rc == 0 and not status.any()
    for b in range(B):
        assert rel_err(got[b], orc.init_demixing(Cx[b], K)) < 1e-12
        old, _ = G.init_demix(Cx[b], K, L.INIT_EYE)
        assert rel_err(got[b], old) < 1e-12
    rng = np.random.default_rng(19)
    W0 = rng.standard_normal((B, F, M, K)) + 1j * rng.standard_normal((B, F, M, K))
    got, status, rc = G.init_demix_grouped(Cx, K, W0=W0)
    assert rc == 0
    for b in range(B):
        assert rel_err(got[b], orc.init_demixing(Cx[b], K, W0=W0[b])) < 1e-11
    # a silent mixture: its J solve is singular, only its status word is set
    Cx[1] = 0
    _, status, rc = G.init_demix_grouped(Cx, K)
    assert rc == 0 and status[0] == 0 and (status[1] & L.STATUS_SINGULAR or K == M)


@pytest.mark.parametrize("M", [1, 2, 3, 4, 6, 8, 11, 16])
def test_eigh(M):
    X = _mix(14, M, 1500, 32)[0]
    Cx = orc.input_covariance(X)
    ev, vec, status = G.eigh(Cx, lapack_phase=True)
    assert status == 0
    want_ev = np.linalg.eigvalsh(Cx)
    assert np.max(np.abs(ev - want_ev) / np.abs(want_ev).max(axis=1, keepdims=True)) < 1e-13
    # A v = lambda v, unitary, LAPACK zgeev phase (largest component real positive)
    resid = Cx @ vec - vec * ev[:, None, :]
    assert np.linalg.norm(resid) / np.linalg.norm(Cx) < 1e-13
    eye = np.conj(vec.swapaxes(1, 2)) @ vec
    assert np.max(np.abs(eye - np.eye(M))) < 1e-13
    idx = np.argmax(np.abs(vec), axis=1)  # (F, M)
    big = np.take_along_axis(vec, idx[:, None, :], axis=1)[:, 0, :]
    assert np.all(big.imag == 0.0) and np.all(big.real > 0.0)


@pytest.mark.parametrize("M,K", [(4, 2), (6, 2), (8, 3)])
def test_init_demix_eig_matches_reference_rule(M, K):
    """overiva.py:103-109: conj of the top-K eigenvectors of np.linalg.eig, ascending."""
    X = _mix(15, M, 1500, 32)[0]
    Cx = orc.input_covariance(X)
    _, vec, _ = G.eigh(Cx, lapack_phase=True)
    got, status = G.init_demix(Cx, K, L.INIT_EIG, evecs=vec)
    assert status == 0
    assert rel_err(got, orc.init_demixing(Cx, K, init_eig=True)) < 1e-11


@pytest.mark.parametrize("M,K,n_samples,dtype,proj_back", [
    (4, 2, 1500, np.complex128, True), (6, 6, 1200, np.complex128, True), (8, 3, 5000, np.complex128, False),
    (16, 4, 2300, np.complex128, True), (5, 1, 1500, np.complex64, True), (3, 2, 4300, np.complex64, True)])
def test_final_demix_and_projection_back(M, K, n_samples, dtype, proj_back):
    B = 2
    X = _mix(16, M, n_samples, 128 if M % 2 == 0 else 32, dtype, B=B)
    _, T, F, _ = X.shape
    rng = np.random.default_rng(17)
    W = rng.standard_normal((B, F, M, M)) + 1j * rng.standard_normal((B, F, M, M))
    X128 = X.astype(np.complex128)
    Cx = np.stack([orc.input_covariance(X128[b]) for b in range(B)])
    Weff = G.projback_filters(W.reshape(B * F, M, M), Cx.reshape(B * F, M, M), K, proj_back)
    Xg = G.grouped(X)
    Y = G.demix_output(Xg, Weff, B, T, F, M, K, G.code_of(dtype))
    assert Y.dtype == dtype and np.all(np.isfinite(Y))
    for b in range(B):
        Xf = np.ascontiguousarray(X128[b].swapaxes(0, 1))
        want = orc.demix(Xf, W[b][:, :, :K]).swapaxes(0, 1)
        if proj_back:
            z = orc.projection_back(want, X128[b][:, :, 0])
            want = want * np.conj(z[None])
        assert rel_err(Y[b], want) < (1e-12 if dtype == np.complex128 else 1e-6)
    # the single-launch version (filters from the grouped W_hat, projection-back scale computed per lane from the grouped
    # covariance): same arithmetic in the same order as the three-kernel sequence above
    Yg = G.demix_output_grouped(Xg, W, Cx, B, T, F, M, K, G.code_of(dtype), proj_back)
    assert Yg.dtype == dtype and rel_err(Yg, Y) < (1e-14 if dtype == np.complex128 else 1e-6)


@pytest.mark.parametrize("M,K,dtype", [(6, 2, np.complex128), (4, 2, np.complex128), (6, 2, np.complex64), (3, 3, np.complex128),
                                       (5, 3, np.complex128), (8, 1, np.complex128), (4, 4, np.complex64), (2, 1, np.complex128),
                                       (6, 3, np.complex128)])
def test_fused_cov_sweep_equals_two_kernels(M, K, dtype, monkeypatch):
    """oiva_cov_ip_update (one pass over X ending in the per-group sweep, covariances in registers) against
    oiva_weighted_cov_ws + oiva_ip_update on the same state: same arithmetic in the same order -> bit-identical W_hat and
    status words; needs >= 4096 bin groups (no frame splitting).  One mixture is silent (zero input -> singular bins): only
    its status word may be set.  Sampled mixtures are checked against the oracle."""
    from overiva_b200.core import DemixPlan

    B, T, F = 2100, 23, 40  # NG = 2 -> 4200 groups, the second group of every mixture ragged (8 of 32 bins)
    tdt = torch.complex128 if dtype == np.complex128 else torch.complex64
    g = torch.Generator(device="cuda").manual_seed(7 + M * 10 + K)
    rdt = torch.float64 if dtype == np.complex128 else torch.float32
    S = torch.randn(B, T, F, K + 1, 2, generator=g, device="cuda", dtype=rdt)
    S = torch.view_as_complex(S * (torch.rand(B, T, 1, K + 1, 1, generator=g, device="cuda", dtype=rdt) ** 3))
    A = torch.view_as_complex(torch.randn(B, 1, F, M, K + 1, 2, generator=g, device="cuda", dtype=rdt))
    N = torch.view_as_complex(torch.randn(B, T, F, M, 2, generator=g, device="cuda", dtype=rdt))
    X = ((A @ S.unsqueeze(-1)).squeeze(-1) + 1e-2 * N).to(tdt).contiguous()
    X[1234] = 0
    out = {}
    for path in ("fused", "two_kernels"):
        if path == "two_kernels":
            monkeypatch.setenv("OIVA_NO_COV_SWEEP", "1")
        else:
            monkeypatch.delenv("OIVA_NO_COV_SWEEP", raising=False)
        plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, X.dtype, X.device)
        plan.load(X)
        plan.init(L.INIT_EYE)
        l0 = plan.launches
        plan.iterate(3)
        covered = K <= 3 and (M <= 5 or (M == 6 and K <= 2) or K == 1)  # (cov_sweep_supported; others fall back)
        assert plan.launches - l0 == (9 if path == "fused" and covered else 12), (path, plan.launches - l0)
        out[path] = (plan.filters().clone(), plan.status_vector().copy())
        del plan
    Wf, sf = out["fused"]
    Wt, st = out["two_kernels"]
    assert np.array_equal(sf, st) and sf[1234] != 0 and not np.any(np.delete(sf, 1234))
    good = torch.ones(B, dtype=torch.bool, device=X.device)
    good[1234] = False
    assert torch.equal(Wf[good], Wt[good])
    for b in (0, 2099):
        _, Wo = orc.overiva(X[b].cpu().numpy().astype(np.complex128), n_src=K, n_iter=3, return_filters=True)
        assert rel_err(Wf[b].cpu().numpy(), Wo) <= (1e-10 if dtype == np.complex128 else 1e-4)


@pytest.mark.parametrize("M,K,n_samples", [(2, 2, 1500), (4, 4, 2300), (6, 6, 1200), (8, 8, 900), (5, 3, 30000), (7, 1, 1500)])
def test_weighted_covariance_per_bin_weights_and_power_output(M, K, n_samples):
    """ILRMA's two kernel-level additions: the covariance with one weight per (source, frame, BIN) -- staged per lane next
    to the samples -- and the per-bin power output of the statistic kernel, against numpy."""
    lib = L.load()
    B = 2
    frame = 16 if n_samples > 20000 else 80  # (30000 samples at frame 16: few bins, many frames -> frame splits)
    X = _mix(41, M, n_samples, frame, np.complex128, B=B)
    _, T, F, _ = X.shape
    NG, Tp = lib.oiva_bin_groups(F), lib.oiva_frame_pitch(T)
    rng = np.random.default_rng(43)
    winv = rng.gamma(1.0, 1.0, size=(B, K, F, T)) + 0.05
    # grouped weights [gi][k][Tp][32], zero on padded bins / frames
    wg = np.zeros((B, NG, K, Tp, 32))
    pad = np.zeros((B, K, NG * 32, Tp))
    pad[:, :, :F, :T] = winv
    wg[:] = pad.reshape(B, K, NG, 32, Tp).transpose(0, 2, 1, 4, 3)
    Xg = G.grouped(X)
    wd = G.to_dev(wg)
    Vg = torch.full((lib.oiva_grouped_cov_bytes(B, F, M, K) // 8,), float("nan"), dtype=torch.float64, device=G.dev())
    ws_bytes = lib.oiva_weighted_cov_scratch_bytes(B, T, F, M, K)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=G.dev())
    L.check(lib.oiva_weighted_cov_binwise(G.P(Xg), G.P(wd), G.P(Vg), G.P(ws) if ws_bytes else None, ws_bytes, B, T, F, M, K,
                                          G.stream()), "oiva_weighted_cov_binwise")
    V = torch.empty((B, F, K, M, M), dtype=torch.complex128, device=G.dev())
    L.check(lib.oiva_unpack_cov(G.P(Vg), G.P(V), B, F, M, K, G.stream()), "oiva_unpack_cov")
    torch.cuda.synchronize()
    V = V.cpu().numpy()
    for b in range(B):
        for k in range(K):
            want = np.einsum("tfm,ft,tfn->fmn", X[b], winv[b, k], np.conj(X[b])) / T
            assert rel_err(V[b, :, k], want) < 1e-12, (b, k)
    # per-bin powers |w_k^H x|^2 and their sums over the bins
    W = rng.standard_normal((B, F, M, K)) + 1j * rng.standard_normal((B, F, M, K))
    Wd = G.to_dev(W.astype(np.complex128))
    part = torch.full((B, NG, K, Tp), np.nan, dtype=torch.float64, device=G.dev())
    Pg = torch.full((B, NG, K, Tp, 32), np.nan, dtype=torch.float64, device=G.dev())
    L.check(lib.oiva_demix_power_full(G.P(Xg), G.P(Wd), K, 0, G.P(part), G.P(Pg), B, T, F, M, K, L.C128, G.stream()),
            "oiva_demix_power_full")
    torch.cuda.synchronize()
    Pw = np.abs(np.einsum("btfm,bfmk->bkft", X, np.conj(W))) ** 2  # (B, K, F, T)
    Pgot = Pg.cpu().numpy().transpose(0, 2, 1, 4, 3).reshape(B, K, NG * 32, Tp)
    assert rel_err(Pgot[:, :, :F, :T], Pw) < 1e-13
    assert not Pgot[:, :, F:, :].any() and not Pgot[:, :, :, T:].any()  # padded bins / frames are written as zeros
    r2 = part.cpu().numpy().sum(axis=1)[:, :, :T]
    assert rel_err(r2, Pw.sum(axis=2)) < 1e-13
