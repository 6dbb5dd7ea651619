"""STFT analysis / synthesis either side of the loop.

CPU: the numpy oracle of the transform pair (perfect reconstruction, agreement with numpy's own rfft framing,
window formulas) and the host-side helpers of ``overiva_b200.stft``.
GPU: the hand-written FFT kernels against the oracle (relative Frobenius error <= 1e-12 in fp64: the kernels use
a radix-2 FFT, numpy uses pocketfft -- results differ at rounding level), the grouped output against the relayout
of the plain output (bit-exact), and the audio-in / audio-out pipeline against oracle analysis -> oracle
overiva -> oracle synthesis (<= 1e-10, the loop's tolerance)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err
from oracle import overiva_oracle as orc
from oracle import stft_oracle as so
from overiva_b200 import _lib as L
from overiva_b200 import stft as gst
from overiva_b200.synth import convolutive_mixture
from overiva_b200.synth import stft as synth_stft

torch = pytest.importorskip("torch")

FFT_TOL = 1e-12
F32_TOL = 1e-5


# ---- CPU ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L_,hop", [(64, 32), (64, 16), (32, 8), (16, 16)])
def test_oracle_perfect_reconstruction(L_, hop):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((L_ * 9 + 5, 3))
    wa = so.hann(L_) if hop < L_ else np.ones(L_)
    ws = so.compute_synthesis_window(wa, hop)
    X = so.analysis(x, L_, hop, win=wa)
    assert X.shape == ((x.shape[0] - L_) // hop + 1, L_ // 2 + 1, 3)
    y = so.synthesis(X, L_, hop, win=ws)
    # fully overlapped region: every hop-shift of the window is present
    lo, hi = L_ - hop, (X.shape[0] - 1) * hop
    assert rel_err(y[lo:hi], x[lo:hi]) < 1e-13


@pytest.mark.parametrize("L_,hop", [(64, 32), (64, 16), (32, 8), (4096, 2048)])
def test_oracle_against_scipy_short_time_fft(L_, hop):
    """An INDEPENDENT implementation pins the restated transform pair (pyroomacoustics itself is absent):
    scipy.signal.ShortTimeFFT with the same window, hop, no scaling and no phase shift computes
    rfft(win * x[p*hop - L/2 : p*hop + L/2]) for slice p, i.e. our frame t = p - (L/2)/hop; its canonical dual window is
    win / sum_p win_p^2 (our ``compute_synthesis_window``) and its inverse is the overlap-add of dual * irfft (our
    ``synthesis``)."""
    from scipy.signal import ShortTimeFFT

    rng = np.random.default_rng(3)
    x = rng.standard_normal(L_ * 7 + 11)
    wa = so.hann(L_)
    sft = ShortTimeFFT(wa, hop, fs=1.0, fft_mode="onesided", scale_to=None, phase_shift=None)
    Z = sft.stft(x)  # (F, slices), slices p_min .. p_max-1
    X = so.analysis(x, L_, hop, win=wa)
    p0 = (L_ // 2) // hop - sft.p_min
    assert rel_err(X, Z[:, p0 : p0 + X.shape[0]].T) < 1e-13
    ws = so.compute_synthesis_window(wa, hop)
    assert rel_err(ws, sft.dual_win) < 1e-13
    # scipy's inverse of ITS transform and our overlap-add of OUR transform agree where every shift is present
    y_scipy = sft.istft(Z, k1=len(x))
    y_ours = so.synthesis(X, L_, hop, win=ws)
    lo, hi = L_ - hop, (X.shape[0] - 1) * hop
    assert rel_err(y_ours[lo:hi], y_scipy[lo:hi]) < 1e-12
    assert rel_err(y_scipy[lo:hi], x[lo:hi]) < 1e-12


def test_oracle_matches_the_generator_stft():
    # overiva_b200.synth.stft (what the golden fixtures were made with) is the same transform
    x = np.random.default_rng(1).standard_normal((700, 2))
    assert rel_err(so.analysis(x, 64, 32, win=so.hann(64)), synth_stft(x, 64, 32)) < 1e-14


def test_oracle_pad_front_is_the_streaming_state_convention():
    # L - hop zeros in front delay the reconstruction by L - hop samples (overiva_oneshot.py:393-401 compares
    # y[framesize//2:] with the clean signals)
    L_, hop = 64, 32
    x = np.random.default_rng(2).standard_normal(640)
    wa = so.hann(L_)
    y = so.synthesis(so.analysis(x, L_, hop, win=wa, pad_front=L_ - hop), L_, hop,
                     win=so.compute_synthesis_window(wa, hop))
    assert rel_err(y[hop : hop + 500], x[:500]) < 1e-13


def test_host_window_helpers_match_the_oracle():
    for n, hop in [(4096, 2048), (64, 16), (48, 16), (30, 7)]:
        assert np.array_equal(gst.hann(n), so.hann(n))
        assert np.allclose(gst.compute_synthesis_window(so.hann(n) + 0.1, hop),
                           so.compute_synthesis_window(so.hann(n) + 0.1, hop), rtol=1e-15, atol=0)


def test_num_frames_and_argument_checks():
    lib = L.load()
    assert lib.oiva_stft_num_frames(240000, 4096, 2048, 0, 0) == 116  # BASELINE cfg 1: 15 s at 16 kHz
    assert lib.oiva_stft_num_frames(240000, 4096, 2048, 2048, 0) == 117
    assert lib.oiva_stft_num_frames(100, 4096, 2048, 0, 0) == 0
    assert gst.num_frames(700, 64, 32) == so.num_frames(700, 64, 32)
    one = C.c_void_p(16)
    assert lib.oiva_stft_twiddles(one, 100, None) == L.C.c_int(-1).value  # not a power of two
    assert b"power of two" in lib.oiva_last_error()
    assert lib.oiva_stft_analysis(one, 0, 0, 1, 1, 100, 0, None, one, one, 0, 1, 1, 1, 64, 0, 0, None) == -1  # hop 0
    assert lib.oiva_stft_synthesis(one, None, one, one, one, 0, 1, 1, 1, 16384, 64, 0, None) == -1  # frame too long
    assert lib.oiva_stft_scratch_bytes(2, 3, 4, 64) == 2 * 3 * 4 * 64 * 8
    # one 48-byte record (w, w^2, w^4) per butterfly phase of every FFT pass after the first: 8 + 64 + 512 at 4096
    assert lib.oiva_stft_twiddle_bytes(4096) == 3 * (8 + 64 + 512) * 16
    assert lib.oiva_stft_twiddle_bytes(64) == 3 * 8 * 16 and lib.oiva_stft_twiddle_bytes(100) == 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gst.analysis(np.zeros((256, 2)), 64, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gst.separate(np.zeros((256, 2)), framesize=64)


# ---- GPU ------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("L_,hop,M,pad", [(8, 4, 1, 0), (64, 32, 3, 0), (64, 16, 2, 48), (256, 100, 4, 0),
                                          (1024, 512, 2, 512), (4096, 2048, 6, 0), (8192, 4096, 1, 0)])
def test_analysis_matches_oracle(L_, hop, M, pad):
    rng = np.random.default_rng(L_ + hop)
    x = rng.standard_normal((L_ * 5 + 37, M))
    win = so.hann(L_)
    ref = so.analysis(x, L_, hop, win=win, pad_front=pad, pad_back=11)
    X = gst.analysis(x, L_, hop, win=win, pad_front=pad, pad_back=11)
    assert X.shape == ref.shape and X.dtype == np.complex128
    assert rel_err(X, ref) <= FFT_TOL
    assert np.all(X[:, 0].imag == 0) and np.all(X[:, -1].imag == 0)


@gpu
def test_analysis_input_kinds_and_strides():
    rng = np.random.default_rng(5)
    xm = rng.standard_normal((3, 1000))  # channel-major memory, passed as its (N, M) transpose like the drivers do
    ref = so.analysis(xm.T, 128, 64, win=so.hann(128))
    assert rel_err(gst.analysis(xm.T, 128, 64, win=so.hann(128)), ref) <= FFT_TOL
    Xt = gst.analysis(torch.from_numpy(xm).cuda().T, 128, 64, win=so.hann(128))
    assert Xt.is_cuda and rel_err(Xt.cpu().numpy(), ref) <= FFT_TOL
    # mono, no window
    assert rel_err(gst.analysis(xm[0], 128, 64), so.analysis(xm[0], 128, 64)) <= FFT_TOL
    # batch
    xb = rng.standard_normal((4, 700, 2))
    Xb = gst.analysis(xb, 64, 32, win=so.hann(64))
    for b in range(4):
        assert rel_err(Xb[b], so.analysis(xb[b], 64, 32, win=so.hann(64))) <= FFT_TOL
    # float32 audio -> complex64 spectra (fp64 arithmetic inside)
    X32 = gst.analysis(xb[0].astype(np.float32), 64, 32, win=so.hann(64))
    assert X32.dtype == np.complex64
    assert rel_err(X32, so.analysis(xb[0].astype(np.float32), 64, 32, win=so.hann(64))) <= F32_TOL
    with pytest.raises(ValueError, match="shorter than one frame"):
        gst.analysis(np.zeros(10), 64, 32)
    with pytest.raises(TypeError):
        gst.analysis(np.zeros(100, dtype=np.complex128), 64, 32)


@gpu
@pytest.mark.parametrize("cdt", [np.complex128, np.complex64])
def test_grouped_output_equals_relayout_of_plain_output(cdt):
    """The analysis kernel writing the loop's grouped layout directly == plain (B,T,F,M) output + oiva_relayout."""
    from gpu_util import P, dev, grouped, stream

    lib = L.load()
    rng = np.random.default_rng(7)
    B, N, M, L_, hop = 3, 1500, 5, 128, 64
    x = rng.standard_normal((B, N, M))
    X = gst.analysis(x, L_, hop, win=so.hann(L_), dtype=cdt)
    T, F = X.shape[1], X.shape[2]
    expect = grouped(X.astype(cdt))
    xd = torch.from_numpy(x).to(dev())
    code = L.C64 if cdt == np.complex64 else L.C128
    got = torch.full((lib.oiva_grouped_bytes(B, T, F, M, code),), 0xFF, dtype=torch.uint8, device=dev())
    win = torch.from_numpy(so.hann(L_)).to(dev())
    tw = torch.empty(lib.oiva_stft_twiddle_bytes(L_) // 16, dtype=torch.complex128, device=dev())
    L.check(lib.oiva_stft_twiddles(P(tw), L_, stream()), "tw")
    L.check(lib.oiva_stft_analysis(P(xd), 0, N * M, M, 1, N, 0, P(win), P(tw), P(got), 1, B, T, M, L_, hop, code,
                                   stream()), "analysis")
    torch.cuda.synchronize()
    assert torch.equal(got, expect)


@gpu
@pytest.mark.parametrize("L_,hop,K", [(8, 4, 1), (64, 32, 2), (64, 16, 3), (256, 100, 2), (4096, 2048, 2),
                                      (8192, 2048, 1)])
def test_synthesis_matches_oracle(L_, hop, K):
    rng = np.random.default_rng(L_ * 3 + hop)
    T, F = 7, L_ // 2 + 1
    Y = rng.standard_normal((T, F, K)) + 1j * rng.standard_normal((T, F, K))  # DC / Nyquist imaginary parts ignored
    ws = so.compute_synthesis_window(so.hann(L_), hop)
    ref = so.synthesis(Y, L_, hop, win=ws)
    y = gst.synthesis(Y, L_, hop, win=ws)
    assert y.shape == ref.shape and y.dtype == np.float64
    assert rel_err(y, ref) <= FFT_TOL
    assert rel_err(gst.synthesis(Y[:, :, 0], L_, hop), so.synthesis(Y[:, :, 0], L_, hop)) <= FFT_TOL
    yb = gst.synthesis(np.stack([Y, 2 * Y]), L_, hop, win=ws)
    assert rel_err(yb[1], 2 * ref) <= FFT_TOL
    y32 = gst.synthesis(Y.astype(np.complex64), L_, hop, win=ws)
    assert y32.dtype == np.float32 and rel_err(y32, ref) <= F32_TOL


@gpu
def test_round_trip_reconstructs_the_signal():
    rng = np.random.default_rng(11)
    L_, hop = 4096, 2048
    x = rng.standard_normal((L_ * 6, 2))
    wa = gst.hann(L_)
    y = gst.synthesis(gst.analysis(x, L_, hop, win=wa), L_, hop, win=gst.compute_synthesis_window(wa, hop))
    assert rel_err(y[hop:-hop], x[hop : y.shape[0] - hop]) <= FFT_TOL


@gpu
@pytest.mark.parametrize("algo,kw", [("overiva", dict(n_src=2)), ("auxiva", {}), ("overiva", dict(n_src=1, model="gauss")),
                                     ("auxiva_pca", dict(n_src=2)), ("ogive", dict(n_iter=30))])
def test_separate_matches_oracle_pipeline(algo, kw):
    """audio in -> audio out on the device == oracle analysis -> oracle algorithm -> oracle synthesis."""
    mix, _ = convolutive_mixture(3, 4, 2, duration=0.35, fs=8000, n_interferers=4, rt60=0.02, env_shape=2.0,
                                 env_block=0.02)
    L_, hop = 64, 32
    wa = so.hann(L_)
    ws = so.compute_synthesis_window(wa, hop)
    X = so.analysis(mix, L_, hop, win=wa)
    n_iter = kw.pop("n_iter", 20)
    if algo == "overiva":
        Yr = orc.overiva(X, n_iter=n_iter, **kw)
    elif algo == "auxiva":
        Yr = orc.overiva(X, n_iter=n_iter)
    elif algo == "auxiva_pca":
        Yr = orc.auxiva_pca(X, n_iter=n_iter, proj_back=True, **kw)
    else:
        Yr = orc.ogive(X, n_iter=n_iter, **kw)
    ref = so.synthesis(Yr, L_, hop, win=ws)
    y = gst.separate(mix, algo=algo, n_iter=n_iter, framesize=L_, **kw)
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 1e-10


@gpu
def test_separate_batch_filters_and_errors():
    mixes = np.stack([convolutive_mixture(s, 3, 2, duration=0.3, fs=8000, n_interferers=3, rt60=0.02, env_shape=2.0,
                                          env_block=0.02)[0] for s in (1, 2, 3)])
    y, W = gst.separate(mixes, n_src=2, framesize=64, return_filters=True)
    assert y.shape[0] == 3 and y.shape[2] == 2 and W.shape == (3, 33, 3, 2)
    for b in range(3):
        yb, Wb = gst.separate(mixes[b], n_src=2, framesize=64, return_filters=True)
        # same arithmetic per mixture; only the frame split of the covariance sums depends on the batch size
        assert rel_err(yb, y[b]) < 1e-11 and rel_err(Wb, W[b]) < 1e-11
    yt = gst.separate(torch.from_numpy(mixes[0]).cuda(), n_src=2, framesize=64)
    assert yt.is_cuda and rel_err(yt.cpu().numpy(), y[0]) < 1e-11
    with pytest.raises(ValueError, match="No such algorithm"):
        gst.separate(mixes[0], algo="fastica")
    # ILRMA through the audio-in / audio-out call (overiva_oneshot.py:331-339): determined, 3 outputs
    np.random.seed(5)
    yi = gst.separate(mixes[0], algo="ilrma", framesize=64, n_iter=5, n_components=2)
    assert yi.shape[1] == 3 and np.all(np.isfinite(yi))
    with pytest.raises(ValueError, match="one mixture at a time"):
        gst.separate(mixes, algo="ogive", framesize=64)
    with pytest.raises(ValueError, match="n_src"):
        gst.separate(mixes[0], n_src=7, framesize=64)


@gpu
def test_separate_batch_host_pipeline_equals_per_mixture_calls():
    mixes = np.stack([convolutive_mixture(s, 3, 2, duration=0.3, fs=8000, n_interferers=3, rt60=0.02, env_shape=2.0,
                                          env_block=0.02)[0] for s in range(5)])
    y = gst.separate_batch(mixes, n_src=2, framesize=64, chunk=2)  # 3 chunks, the last one partial
    assert y.shape == (5, (mixes.shape[1] - 64) // 32 * 32 + 64, 2) and y.dtype == np.float64
    for b in range(5):
        assert rel_err(y[b], gst.separate(mixes[b], n_src=2, framesize=64)) < 1e-11
    pinned = torch.from_numpy(mixes).pin_memory()
    out = torch.empty(y.shape, dtype=torch.float64).pin_memory()
    yt = gst.separate_batch(pinned, n_src=2, framesize=64, chunk=4, out=out)
    assert yt is out and rel_err(out.numpy(), y) < 1e-11
    y32 = gst.separate_batch(mixes.astype(np.float32), n_src=2, framesize=64, dtype=torch.complex64)
    assert y32.dtype == np.float32 and rel_err(y32, y) < 1e-3
