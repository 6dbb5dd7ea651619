"""GPU parity tests of the drop-in entry points (overiva / auxiva / auxiva_pca / ogive / overiva_batch)
against the golden vectors produced by the unmodified reference and against the numpy oracle.

Tolerances (BASELINE.json north_star): relative Frobenius error <= 1e-10 on W and Y in the fp64 mode
after 20 iterations; <= 1e-3 in the fp32-storage / fp64-solve mode."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle import overiva_oracle as orc
from overiva_b200.synth import convolutive_mixture, small_test_mixture, stft

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if torch.cuda.is_available():
    import overiva_b200 as ob

FP64_TOL = 1e-10
FP32_TOL = 1e-3


def _run(case):
    kw = dict(case["kwargs"])
    if "W0" in case:
        kw["W0"] = case["W0"]
    fn = case["fn"]
    if fn == "overiva":
        return ob.overiva(case["X"], return_filters=True, **kw)
    if fn == "auxiva_pca":
        return ob.auxiva_pca(case["X"], **kw), None
    if fn == "ogive":
        return ob.ogive(case["X"], return_filters=True, **kw)
    raise ValueError(fn)


@pytest.fixture(params=["resident", "kernel_loop"])
def loop_path(request, monkeypatch):
    """Short inputs run the persistent single-launch loop (csrc/resident.cuh); OIVA_NO_RESIDENT=1 forces the
    kernel-per-step loop (CUDA graph / eager) for the same shapes, so both paths see every case."""
    if request.param == "kernel_loop":
        monkeypatch.setenv("OIVA_NO_RESIDENT", "1")
    else:
        monkeypatch.delenv("OIVA_NO_RESIDENT", raising=False)
    if torch.cuda.is_available():
        ob.clear_plan_cache()
    return request.param


@pytest.mark.parametrize("name", golden_names())
def test_against_reference_golden(name, loop_path):
    case = load_golden(name)
    Y, W = _run(case)
    c64 = case["X"].dtype == np.complex64
    tol = FP32_TOL if c64 else max(FP64_TOL, 1e2 * case["sens"])
    assert Y.shape == case["Y"].shape and Y.dtype == case["Y"].dtype
    assert np.all(np.isfinite(Y))
    assert rel_err(Y, case["Y"]) <= tol, "Y"
    if W is not None:
        assert W.shape == case["W"].shape and W.dtype == case["W"].dtype
        assert rel_err(W, case["W"]) <= tol, "W"


def test_fp32_storage_mode_vs_fp64_reference():
    """fp32 storage of X/Y with fp64 covariance + solve: within 1e-3 of the reference run in complex128."""
    case = load_golden("overiva_laplace_m6k2")
    X = case["X"]
    Y, W = ob.overiva(X.astype(np.complex64), return_filters=True, **case["kwargs"])
    assert Y.dtype == np.complex64
    assert rel_err(Y.astype(np.complex128), case["Y"]) <= FP32_TOL
    assert rel_err(W.astype(np.complex128), case["W"]) <= FP32_TOL


# BASELINE.json configs at their real STFT shape (F = 2049), on convolutive mixtures, against the oracle
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3"])
def test_baseline_configs_full_bins(cfg, loop_path):
    if cfg == "cfg1":  # overiva -m 4 -s 2 -n 20, 15 s @ 16 kHz
        mix, _ = convolutive_mixture(101, 4, 2, duration=15.0)
        kw = dict(n_src=2, n_iter=20, model="laplace")
    elif cfg == "cfg2":  # auxiva determined M = K = 6
        mix, _ = convolutive_mixture(102, 6, 2, duration=15.0)
        kw = dict(n_iter=20, model="laplace")
    else:  # overiva M=8 K=2 gauss init_eig, the full 60 s mixture of BASELINE config 3 (T = 467)
        if loop_path == "resident":
            pytest.skip("config 3 does not fit the resident loop (122 MB of samples): one path only")
        mix, _ = convolutive_mixture(103, 8, 2, duration=60.0)
        kw = dict(n_src=2, n_iter=20, model="gauss", init_eig=True)
    X = stft(mix)
    assert X.shape[1] == 2049 and (cfg != "cfg3" or X.shape[0] == 467)
    Yo, Wo = orc.overiva(X, return_filters=True, **kw)
    Y, W = ob.overiva(X, return_filters=True, **kw)
    assert rel_err(Y, Yo) <= FP64_TOL
    assert rel_err(W, Wo) <= FP64_TOL


@pytest.mark.parametrize("M,K,model,dtype,n_samples,frame", [
    (4, 2, "laplace", np.complex128, 30000, 512), (6, 6, "laplace", np.complex128, 20000, 512),
    (3, 1, "gauss", np.complex128, 9000, 128), (8, 4, "laplace", np.complex128, 16000, 256),
    (5, 5, "gauss", np.complex64, 12000, 256), (2, 2, "laplace", np.complex128, 700, 64),
    (6, 2, "laplace", np.complex64, 40000, 1024), (7, 3, "gauss", np.complex128, 8000, 64)])
def test_resident_loop_equals_kernel_loop(M, K, model, dtype, n_samples, frame, monkeypatch):
    """The persistent single-launch loop against the kernel-per-step loop on the same input: same statistic and sweep
    arithmetic, covariance frames summed in different sub-ranges -> agreement to rounding (<= 1e-11 after 12 epochs),
    for one mixture and for a small batch, both storage types; and the resident loop is deterministic."""
    from overiva_b200 import _lib as L
    from overiva_b200.core import DemixPlan

    Xs = np.stack([small_test_mixture(40 + b, M, min(K, 2), n_samples=n_samples, frame=frame, hop=frame // 2)
                   for b in range(2)]).astype(dtype)
    for X in (Xs[:1], Xs):
        Xd = torch.from_numpy(X).cuda()
        B, T, F, _ = X.shape
        out = {}
        for path in ("resident", "kernel_loop", "resident_again"):
            if path == "kernel_loop":
                monkeypatch.setenv("OIVA_NO_RESIDENT", "1")
            else:
                monkeypatch.delenv("OIVA_NO_RESIDENT", raising=False)
            plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE if model == "laplace" else L.MODEL_GAUSS, Xd.dtype, Xd.device)
            plan.load(Xd)
            plan.init(L.INIT_EYE)
            l0 = plan.launches
            plan.iterate(12)
            n_launch = plan.launches - l0
            assert (n_launch == 1) == (path != "kernel_loop"), (path, n_launch)  # the resident loop really ran
            out[path] = (plan.output(True).cpu().numpy(), plan.filters().cpu().numpy())
            plan.raise_on_failure()
        tol = 1e-11 if dtype == np.complex128 else 1e-6
        assert rel_err(out["resident"][0], out["kernel_loop"][0]) <= tol
        assert rel_err(out["resident"][1], out["kernel_loop"][1]) <= tol
        assert np.array_equal(out["resident"][0], out["resident_again"][0])
        assert np.array_equal(out["resident"][1], out["resident_again"][1])


@pytest.mark.parametrize("M,K,n_samples,frame,nb", [(4, 2, 30000, 512, 1), (6, 6, 20000, 512, 1), (4, 2, 1500, 64, 3)])
def test_resident_loop_synchronisation_modes_agree(M, K, n_samples, frame, nb, monkeypatch):
    """The resident loop's hand-over variants -- thread-block clusters or flags in global memory inside a bin group,
    sign-tagged statistic words or two grid barriers per epoch -- move the same numbers in the same order: bit-identical
    demixing matrices (the fallbacks stay usable: OIVA_RES_CLUSTER=0, OIVA_RES_TAGGED=0)."""
    from overiva_b200 import _lib as L
    from overiva_b200.core import DemixPlan

    X = np.stack([small_test_mixture(70 + b, M, 2, n_samples=n_samples, frame=frame, hop=frame // 2) for b in range(nb)])
    Xd = torch.from_numpy(X).cuda()
    B, T, F, _ = X.shape
    out = {}
    for cluster in ("1", "0"):
        for tagged in ("1", "0"):
            monkeypatch.setenv("OIVA_RES_CLUSTER", cluster)
            monkeypatch.setenv("OIVA_RES_TAGGED", tagged)
            plan = DemixPlan(B, T, F, M, K, L.MODEL_LAPLACE, Xd.dtype, Xd.device)
            plan.load(Xd)
            plan.init(L.INIT_EYE)
            l0 = plan.launches
            plan.iterate(9)  # (odd: the last epoch's parity differs from the first's)
            assert plan.launches - l0 == 1
            plan.iterate(4)  # a second launch on the words the first one left behind
            plan.raise_on_failure()
            out[cluster, tagged] = plan.filters().cpu().numpy()
    for key, W in out.items():
        assert np.array_equal(W, out["1", "1"]), key


@pytest.mark.parametrize("M,n_samples,frame", [(6, 20000, 512), (4, 30000, 512), (3, 9000, 128), (5, 12000, 256)])
def test_resident_loop_tracked_inverse_sweep(M, n_samples, frame, monkeypatch):
    """Determined shapes in the resident loop: w_s = V_s^-1 (W^-H e_s) with W^-1 carried along by rank-one updates
    (one Cholesky per source) against the pair sweep that forms W^H V_s and solves it with a pivoted LU
    (OIVA_RES_TRACKED=0) -- the same numbers up to rounding, and both within the tolerance of the oracle."""
    X = small_test_mixture(90 + M, M, 2, n_samples=n_samples, frame=frame, hop=frame // 2)
    Yo, Wo = orc.overiva(X, n_iter=15, return_filters=True)
    out = {}
    for tracked in ("1", "0"):
        monkeypatch.setenv("OIVA_RES_TRACKED", tracked)
        out[tracked] = ob.overiva(X, n_iter=15, return_filters=True)
        assert rel_err(out[tracked][0], Yo) <= FP64_TOL and rel_err(out[tracked][1], Wo) <= FP64_TOL
    assert rel_err(out["1"][1], out["0"][1]) <= 1e-11
    bad = X.copy()
    bad[:, :, M - 1] = bad[:, :, 0]  # rank-deficient: the reference's np.linalg.solve raises
    monkeypatch.setenv("OIVA_RES_TRACKED", "1")
    with pytest.raises(np.linalg.LinAlgError):
        ob.overiva(bad, n_iter=6)


def test_resident_loop_singular_mixture_does_not_stall():
    """A rank-deficient mixture fills the statistic with NaNs; the sign-tagged words must still carry the epoch's parity
    (a NaN that went through an arithmetic instruction comes back as the canonical NaN, sign bit set): LinAlgError like
    the reference, not a stalled loop."""
    X = small_test_mixture(81, 4, 2, n_samples=30000, frame=512, hop=256)
    X[:, :, 3] = X[:, :, 0]
    with pytest.raises(np.linalg.LinAlgError):
        ob.overiva(X, n_src=2, n_iter=7)


def test_batch_equals_single_and_is_deterministic():
    Xs = np.stack([small_test_mixture(200 + b, 6, 2, n_samples=2000, frame=64, hop=32) for b in range(5)])
    Yb, Wb = ob.overiva_batch(Xs, n_src=2, n_iter=10, return_filters=True)
    Yb2 = ob.overiva_batch(Xs, n_src=2, n_iter=10)
    # bit-reproducible at every size: large batches never split a bin group; small problems split its frames over
    # several teams and add the per-split partial covariances in a fixed order
    assert np.array_equal(Yb2, Yb)
    for b in range(5):
        Y1, W1 = ob.overiva(Xs[b], n_src=2, n_iter=10, return_filters=True)
        Yo, Wo = orc.overiva(Xs[b], n_src=2, n_iter=10, return_filters=True)
        assert rel_err(Yb[b], Yo) <= FP64_TOL and rel_err(Wb[b], Wo) <= FP64_TOL
        assert rel_err(Yb[b], Y1) <= 1e-12


def test_cuda_tensor_in_cuda_tensor_out():
    X = small_test_mixture(210, 4, 2)
    Xd = torch.from_numpy(X).cuda()
    Y = ob.overiva(Xd, n_src=2, n_iter=5)
    assert isinstance(Y, torch.Tensor) and Y.is_cuda and Y.dtype == torch.complex128
    assert rel_err(Y.cpu().numpy(), orc.overiva(X, n_src=2, n_iter=5)) <= FP64_TOL
    assert torch.equal(Xd.cpu(), torch.from_numpy(X))  # input not mutated


def test_callback_cadence_and_content():
    X = small_test_mixture(211, 3, 2, n_samples=1200, frame=32, hop=16)
    got, want = [], []
    ob.overiva(X, n_src=2, n_iter=25, callback=lambda Y: got.append(np.array(Y)))
    orc.overiva(X, n_src=2, n_iter=25, callback=lambda Y: want.append(Y.copy()))
    assert len(got) == len(want) == 3  # epochs 0, 10, 20 (overiva.py:142)
    for a, b in zip(got, want):
        assert rel_err(a, b) <= FP64_TOL


def test_errors_mirror_the_reference():
    X = small_test_mixture(212, 3, 2, n_samples=800, frame=32, hop=16)
    with pytest.raises(ValueError):
        ob.overiva(X, n_src=4)  # more sources than channels
    with pytest.raises(KeyError):
        ob.auxiva_pca(X, n_src=2, n_iter=3)  # proj_back is mandatory in auxiva_pca (auxiva_pca.py:86)
    # rank-deficient input -> LinAlgError("Singular matrix"), as numpy raises from overiva.py:98/182
    Xs = X.copy()
    Xs[:, :, 2] = Xs[:, :, 0]
    with pytest.raises(np.linalg.LinAlgError):
        ob.overiva(Xs, n_src=2, n_iter=3)


def test_sdr_sir_within_0p1_db_of_oracle():
    from overiva_b200.metrics import bss_eval
    from overiva_b200.synth import istft

    mix, images = convolutive_mixture(300, 4, 2, duration=6.0, n_interferers=4)
    X = stft(mix, 1024, 512)
    refs = images[:, :, 0]
    out = {}
    for name, fn, x in (("oracle", orc.overiva, X), ("gpu64", ob.overiva, X),
                        ("gpu32", ob.overiva, X.astype(np.complex64))):
        Y = np.asarray(fn(x, n_src=2, n_iter=20)).astype(np.complex128)
        y = istft(Y, 1024, 512)
        sdr, sir, _ = bss_eval(refs, y.T)
        out[name] = (sdr, sir)
    for name in ("gpu64", "gpu32"):
        assert np.max(np.abs(out[name][0] - out["oracle"][0])) < 0.1
        assert np.max(np.abs(out[name][1] - out["oracle"][1])) < 0.1
    assert out["oracle"][1].mean() > 3.0  # the algorithm does separate on this synthetic mixture


def test_config5_full_length_reduced_bins():
    """BASELINE config 5 at its full LENGTH (T = 14061 frames, M = 16, K = 4) on 64 bins, 3 iterations, against the
    oracle: the frame-split x tiled-covariance x many-group-statistic combination at the real frame count."""
    from overiva_b200.synth import stft_domain_mixture

    X = stft_domain_mixture(77, 14061, 64, 16, 4, n_interferers=16)  # >= M sources: well conditioned (sens. 6e-14)
    Yo, Wo = orc.overiva(X, n_src=4, n_iter=3, return_filters=True)
    Y, W = ob.overiva(X, n_src=4, n_iter=3, return_filters=True)
    assert rel_err(Y, Yo) <= FP64_TOL and rel_err(W, Wo) <= FP64_TOL


def test_host_batch_pipeline_equals_device_batch():
    """Host batches larger than `chunk` are streamed through the GPU in overlapped chunks: same results as the
    one-piece device path, including a ragged last chunk, W0 slicing and the accumulated failure status."""
    Xs = np.stack([small_test_mixture(500 + b, 4, 2, n_samples=1500, frame=64, hop=32) for b in range(7)])
    rng = np.random.default_rng(3)
    B, T, F, M = Xs.shape
    W0 = np.zeros((B, F, M, 2), dtype=np.complex128)
    W0[:, :, :2, :] = np.eye(2)
    W0 += 0.05 * (rng.standard_normal(W0.shape) + 1j * rng.standard_normal(W0.shape))
    Yd, Wd = ob.overiva_batch(torch.from_numpy(Xs).cuda(), n_src=2, n_iter=6, W0=W0, return_filters=True)
    Yh, Wh = ob.overiva_batch(Xs, n_src=2, n_iter=6, W0=W0, return_filters=True, chunk=3)  # chunks of 3, 3, 1
    assert isinstance(Yh, np.ndarray) and Yh.shape == (B, T, F, 2)
    assert rel_err(Yh, Yd.cpu().numpy()) <= 1e-12 and rel_err(Wh, Wd.cpu().numpy()) <= 1e-12
    out = torch.empty((B, T, F, 2), dtype=torch.complex128, pin_memory=True)
    Yp = ob.overiva_batch(torch.from_numpy(Xs).pin_memory(), n_src=2, n_iter=6, W0=W0, chunk=2, out=out)
    assert Yp is out and rel_err(out.numpy(), Yh) <= 1e-12
    bad = Xs.copy()
    bad[4, :, :, 3] = bad[4, :, :, 0]  # one rank-deficient mixture in the middle chunk
    with pytest.raises(np.linalg.LinAlgError, match="mixture 4 of 7"):
        ob.overiva_batch(bad, n_src=2, n_iter=3, chunk=3)
    # per-mixture status: only the offending mixture is flagged, the others are separated as usual
    # (the reference's sweep records NaN for the failing task only, overiva_sim.py:334-350)
    for kw in (dict(chunk=3), dict(chunk=64)):  # host pipeline / one-piece path
        Yb, status = ob.overiva_batch(bad, n_src=2, n_iter=6, W0=W0, return_status=True, **kw)
        assert status.shape == (B,) and status[4] & 1 and not np.any(np.delete(status, 4))
        good = np.delete(np.arange(B), 4)
        assert rel_err(Yb[good], Yh[good]) <= 1e-12


def test_large_batch():
    """>= 1184 bin groups (no frame splitting anywhere); results vs the oracle"""
    B = 37
    Xs = np.stack([small_test_mixture(600 + b, 4, 2, n_samples=20000, frame=2048, hop=1024) for b in range(B)])
    assert Xs.shape[2] == 1025 and B * 33 >= 1184
    Y, W = ob.overiva_batch(torch.from_numpy(Xs).cuda(), n_src=2, n_iter=7, return_filters=True)
    Y, W = Y.cpu().numpy(), W.cpu().numpy()
    for b in (0, 17, 36):
        Yo, Wo = orc.overiva(Xs[b], n_src=2, n_iter=7, return_filters=True)
        assert rel_err(Y[b], Yo) <= FP64_TOL and rel_err(W[b], Wo) <= FP64_TOL
