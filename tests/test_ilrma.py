"""ILRMA: the numpy restatement on the CPU, the CUDA path against it on the GPU.
pyroomacoustics is absent, so the oracle is "parity unpinned" (oracle/ilrma_oracle.py); what is pinned here is that the
CUDA path computes exactly what the restatement computes, and that the restatement separates."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ilrma_oracle as ilo
from overiva_b200 import metrics
from overiva_b200.synth import convolutive_mixture, small_test_mixture, stft


def _factors(seed, K, F, T, L):
    rng = np.random.default_rng(seed)
    return 0.1 + 0.9 * rng.random((K, F, L)), 0.1 + 0.9 * rng.random((K, T, L))


def test_oracle_shapes_determinism_and_errors():
    X = small_test_mixture(5, 3, 3, n_samples=1500, frame=64, hop=32, n_interferers=0)
    T, F, M = X.shape
    T0, V0 = _factors(1, M, F, T, 2)
    Y1, W1 = ilo.ilrma(X, n_iter=5, proj_back=True, T0=T0, V0=V0, return_filters=True)
    Y2 = ilo.ilrma(X, n_iter=5, proj_back=True, T0=T0, V0=V0)
    assert Y1.shape == (T, F, M) and W1.shape == (F, M, M) and Y1.dtype == X.dtype
    assert np.array_equal(Y1, Y2) and np.all(np.isfinite(Y1))
    # default factors come from numpy's global generator, T first then V
    np.random.seed(7)
    Ya = ilo.ilrma(X, n_iter=3)
    np.random.seed(7)
    Tr = 0.1 + 0.9 * np.random.rand(M, F, 2)
    Vr = 0.1 + 0.9 * np.random.rand(M, T, 2)
    assert np.array_equal(Ya, ilo.ilrma(X, n_iter=3, T0=Tr, V0=Vr))
    with pytest.raises(ValueError):
        ilo.ilrma(X, n_src=2)
    # n_iter = 0: the demix with the initial W
    assert np.array_equal(ilo.ilrma(X, n_iter=0), X)


def test_oracle_scale_normalisation_and_callback():
    X = small_test_mixture(6, 2, 2, n_samples=2500, frame=64, hop=32, n_interferers=0)
    T, F, M = X.shape
    T0, V0 = _factors(2, M, F, T, 3)
    seen = []
    Y, W = ilo.ilrma(X, n_iter=25, n_components=3, T0=T0, V0=V0, return_filters=True, callback=lambda y: seen.append(y.copy()))
    assert len(seen) == 3 and np.array_equal(seen[0], X)  # epochs 0, 10, 20; the first one sees the initial demix
    # after the last epoch the demixed outputs of the returned W have unit mean power per source
    Yw = np.einsum("tfm,fmk->tfk", X, np.conj(W))
    assert np.allclose(np.mean(np.abs(Yw) ** 2, axis=(0, 1)), 1.0, rtol=1e-10)


def test_oracle_separates_a_determined_mixture():
    """two sources, two microphones, short reverberation: ILRMA must beat the unprocessed mixture by a wide margin"""
    fs, frame, hop = 8000, 256, 128
    mix, refs = convolutive_mixture(11, 2, 2, duration=4.0, fs=fs, n_interferers=0, rt60=0.03)
    X = stft(mix, frame, hop)
    T, F, M = X.shape
    T0, V0 = _factors(3, M, F, T, 2)
    Y = ilo.ilrma(X, n_iter=60, proj_back=True, T0=T0, V0=V0)
    from overiva_b200.stft import hann, compute_synthesis_window
    from oracle import stft_oracle as so

    wa = hann(frame)
    ws = compute_synthesis_window(wa, hop)
    y = so.synthesis(Y, frame, hop, ws)
    n = min(y.shape[0], refs.shape[1])
    sdr, sir, _ = metrics.bss_eval(refs[:2, :n, 0], y[:n].T)
    sdr0, sir0, _ = metrics.bss_eval(refs[:2, :n, 0], mix[:n, :2].T)
    assert np.mean(sir) > np.mean(sir0) + 6.0, (sir, sir0)


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("M,L,proj_back,dtype,n_iter", [(2, 2, True, np.complex128, 15), (3, 2, False, np.complex128, 12),
                                                         (4, 3, True, np.complex128, 12), (6, 2, True, np.complex128, 10),
                                                         (8, 2, True, np.complex128, 8), (3, 4, True, np.complex64, 10),
                                                         (1, 2, False, np.complex128, 5)])
def test_cuda_ilrma_matches_the_oracle(M, L, proj_back, dtype, n_iter):
    import overiva_b200 as ob

    X = small_test_mixture(20 + M, M, min(M, 2), n_samples=3000, frame=80, hop=40, n_interferers=max(M - 2, 0)).astype(dtype)
    T, F, _ = X.shape  # F = 41: a ragged second bin group
    T0, V0 = _factors(4, M, F, T, L)
    Yo, Wo = ilo.ilrma(X, n_iter=n_iter, proj_back=proj_back, n_components=L, T0=T0, V0=V0, return_filters=True)
    Y, W = ob.ilrma(X, n_iter=n_iter, proj_back=proj_back, n_components=L, T0=T0, V0=V0, return_filters=True)
    assert Y.shape == Yo.shape and Y.dtype == X.dtype and W.shape == Wo.shape
    tol = 1e-8 if dtype == np.complex128 else 2e-4
    assert rel_err(Y, Yo) <= tol and rel_err(W, Wo) <= tol, (rel_err(Y, Yo), rel_err(W, Wo))


@gpu
def test_cuda_ilrma_api_behaviour():
    import torch

    import overiva_b200 as ob

    X = small_test_mixture(31, 3, 2, n_samples=2600, frame=64, hop=32, n_interferers=1)
    T, F, M = X.shape
    T0, V0 = _factors(5, M, F, T, 2)
    # callback cadence and content (epochs 0, 10, 20), with and without projection back
    for pb in (True, False):
        got, want = [], []
        Y = ob.ilrma(X, n_iter=21, proj_back=pb, T0=T0, V0=V0, callback=lambda y: got.append(np.array(y)))
        Yo = ilo.ilrma(X, n_iter=21, proj_back=pb, T0=T0, V0=V0, callback=lambda y: want.append(np.array(y)))
        assert len(got) == len(want) == 3
        for a, b in zip(got, want):
            assert rel_err(a, b) <= 1e-8
        assert rel_err(Y, Yo) <= 1e-8
    # n_iter = 0, W0, CUDA tensor in -> CUDA tensor out, global-generator default factors
    assert rel_err(ob.ilrma(X, n_iter=0), X) <= 1e-15
    rng = np.random.default_rng(9)
    W0 = np.eye(M)[None] + 0.1 * (rng.standard_normal((F, M, M)) + 1j * rng.standard_normal((F, M, M)))
    assert rel_err(ob.ilrma(X, n_iter=4, W0=W0, T0=T0, V0=V0), ilo.ilrma(X, n_iter=4, W0=W0, T0=T0, V0=V0)) <= 1e-9
    Yd = ob.ilrma(torch.from_numpy(X).cuda(), n_iter=3, T0=T0, V0=V0)
    assert isinstance(Yd, torch.Tensor) and Yd.is_cuda and Yd.dtype == torch.complex128
    np.random.seed(3)
    Ya = ob.ilrma(X, n_iter=3)
    np.random.seed(3)
    assert rel_err(Ya, ilo.ilrma(X, n_iter=3)) <= 1e-9
    with pytest.raises(ValueError):
        ob.ilrma(X, n_src=2)
    with pytest.raises(ValueError):
        ob.ilrma(np.zeros((10, 5, 9), dtype=np.complex128))
    with pytest.raises(ValueError):
        ob.ilrma(X, n_components=9)
