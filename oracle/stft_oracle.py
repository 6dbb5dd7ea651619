"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the STFT pair the reference's drivers wrap around the loop.

The drivers call pyroomacoustics (pinned ==0.1.23, ``environment.yml:13``; third-party, NOT in the reference
tree and not installed here):

* ``win_a = pra.hann(framesize)``                                         overiva_oneshot.py:157, overiva_sim.py:99
* ``win_s = pra.transform.compute_synthesis_window(win_a, framesize//2)`` overiva_oneshot.py:158, overiva_sim.py:100
* ``X = pra.transform.analysis(mix.T, framesize, framesize//2, win=win_a)``  overiva_oneshot.py:293-295
* ``y = pra.transform.synthesis(Y, framesize, framesize//2, win=win_s)``     overiva_oneshot.py:371-379

The package's sources are not available offline, so the functions below restate its published behaviour (DFT
analysis ``X[t] = rfft(win_a * frame_t)`` without scaling, synthesis = overlap-add of ``win_s * irfft(Y[t])``,
synthesis window = analysis window divided by the sum of its squared hop-shifted copies) and they define what the
CUDA kernels in ``overiva_b200/csrc/stft.cu`` are checked against.  PINNING: against pyroomacoustics itself the
parity is unpinned (absent); the transform pair is pinned against an independent implementation instead --
``scipy.signal.ShortTimeFFT`` (scipy 1.18) with the same window / hop / no scaling gives the same spectra, the same
dual (synthesis) window and the same reconstruction to 1e-13 (``tests/test_stft.py::
test_oracle_against_scipy_short_time_fft``).  Framing follows
SURVEY.md section 8(d): no padding, ``T = (N - L)//hop + 1``; ``pad_front`` zeros may be prepended (the
``L - hop`` state buffer of a streaming STFT, which is what makes the reference compare ``y[framesize//2:]``
with the clean signals, overiva_oneshot.py:393-401).

Only ``tests/`` imports this module.
"""
import numpy as np


def hann(n):
    """Periodic ("asymmetric") Hann window of length n: 0.5 (1 - cos(2 pi i / n))."""
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(n) / n))


def compute_synthesis_window(win_a, hop):
    """win_s[i] = win_a[i] / sum_j win_a[i + j hop]^2 over every hop-shift that still overlaps sample i:
    with it, overlap-add of win_s * (win_a * x) reconstructs x wherever all shifts are present."""
    win_a = np.asarray(win_a, dtype=np.float64)
    L = win_a.shape[0]
    norm = np.zeros(L)
    n = 0
    while n - hop > -L:
        n -= hop
    while n < L:
        if n == 0:
            norm += win_a**2
        elif n < 0:
            norm[: n + L] += win_a[-n - L :] ** 2
        else:
            norm[n:] += win_a[:-n] ** 2
        n += hop
    return win_a / norm


def num_frames(n_samples, L, hop, pad_front=0, pad_back=0):
    n = n_samples + pad_front + pad_back
    return 0 if n < L else (n - L) // hop + 1


def analysis(x, L, hop, win=None, pad_front=0, pad_back=0):
    """x (N,) or (N, M) real -> X (T, F) or (T, F, M) complex128."""
    x = np.asarray(x, dtype=np.float64)
    mono = x.ndim == 1
    if mono:
        x = x[:, None]
    xp = np.concatenate([np.zeros((pad_front, x.shape[1])), x, np.zeros((pad_back, x.shape[1]))])
    T = num_frames(x.shape[0], L, hop, pad_front, pad_back)
    X = np.empty((T, L // 2 + 1, x.shape[1]), dtype=np.complex128)
    w = np.ones(L) if win is None else np.asarray(win, dtype=np.float64)
    for t in range(T):
        X[t] = np.fft.rfft(xp[t * hop : t * hop + L] * w[:, None], axis=0)
    return X[:, :, 0] if mono else X


def synthesis(X, L, hop, win=None):
    """X (T, F) or (T, F, K) -> y ((T-1) hop + L,) or (..., K) real: overlap-add of win * irfft(X[t])."""
    X = np.asarray(X)
    mono = X.ndim == 2
    if mono:
        X = X[:, :, None]
    T, _, K = X.shape
    w = np.ones(L) if win is None else np.asarray(win, dtype=np.float64)
    y = np.zeros(((T - 1) * hop + L, K))
    for t in range(T):
        y[t * hop : t * hop + L] += np.fft.irfft(X[t], n=L, axis=0) * w[:, None]
    return y[:, 0] if mono else y
