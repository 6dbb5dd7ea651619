"""Seeded synthetic inputs for the OverIVA path (CMU ARCTIC / pyroomacoustics are not available
offline -- BASELINE.json ``north_star``; recipe in SURVEY.md section 8d, which mirrors the mixing
rules of the reference's sweep driver ``overiva_sim.py:114-192``).

Two generators:

* :func:`convolutive_mixture` -- time-domain: Laplacian sources with a speech-like block envelope,
  random exponentially decaying RIRs (RT60 0.3 s), 10 interferers at SINR 10 dB, sensor noise at
  SNR 60 dB; returns the mixture and the per-source images so SDR/SIR can be evaluated.
* :func:`stft_domain_mixture` -- draws X directly in the STFT domain (complex super-Gaussian
  sources times a random per-bin mixing matrix plus a noise floor); used for large throughput
  runs where minutes of host-side convolution would dominate set-up.

Plus the STFT pair used by the reference's drivers (frame 4096, hop 2048, Hann analysis window:
``overiva_oneshot.py:156-158,293-295``), here without padding: ``T = (N - frame)//hop + 1``.
"""
from __future__ import annotations

import numpy as np

FRAME = 4096
HOP = 2048


def n_frames(n_samples: int, frame: int = FRAME, hop: int = HOP) -> int:
    return (n_samples - frame) // hop + 1


def stft(x, frame: int = FRAME, hop: int = HOP):
    """x: (N, M) real -> X: (T, F, M) complex128, Hann analysis window, no padding."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    T = n_frames(x.shape[0], frame, hop)
    win = np.hanning(frame + 1)[:-1]  # periodic Hann (perfect reconstruction at 50 % overlap)
    idx = np.arange(frame)[None, :] + hop * np.arange(T)[:, None]
    frames = x[idx, :] * win[None, :, None]  # (T, frame, M)
    return np.fft.rfft(frames, axis=1).astype(np.complex128)


def istft(X, frame: int = FRAME, hop: int = HOP):
    """X: (T, F, K) -> (N, K); overlap-add with a rectangular synthesis window, which is the
    matched synthesis for a periodic Hann at 50 % overlap (sum of shifted windows == 1)."""
    X = np.asarray(X)
    T = X.shape[0]
    frames = np.fft.irfft(X, n=frame, axis=1)  # (T, frame, K)
    out = np.zeros(((T - 1) * hop + frame, X.shape[2]))
    for t in range(T):
        out[t * hop : t * hop + frame] += frames[t]
    return out


def _speechlike(rng, n, fs, env_shape=0.5, env_block=0.25):
    """i.i.d. Laplace(0,1) samples times a piece-wise constant (250 ms) Gamma(0.5) envelope."""
    blk = max(1, int(env_block * fs))
    env = rng.gamma(env_shape, 1.0 / env_shape, size=n // blk + 1)
    env = np.repeat(env, blk)[:n]
    return rng.laplace(0.0, 1.0, size=n) * env


def _rir(rng, fs, rt60=0.3, length=None):
    if length is None:
        length = int(rt60 * fs)
    n = np.arange(length)
    h = rng.standard_normal(length) * np.exp(-6.9 * n / (rt60 * fs)) * 0.1
    d = int(rng.integers(1, 41))
    h[:d] = 0.0
    h[d] = 1.0  # dominant direct-path tap
    return h


def convolutive_mixture(
    seed,
    n_mics,
    n_targets,
    duration=15.0,
    fs=16000,
    n_interferers=10,
    sinr_db=10.0,
    snr_db=60.0,
    rt60=0.3,
    env_shape=0.5,
    env_block=0.25,
):
    """Returns ``(mix (N, M), images (n_targets, N, M))`` -- ``images[k]`` is target k as recorded
    at every microphone (the quantity SDR/SIR are evaluated against, as the reference's drivers do
    with ``premix``: ``overiva_sim.py:163-200``)."""
    rng = np.random.default_rng(seed)
    n = int(duration * fs)
    n_src = n_targets + n_interferers
    L = int(rt60 * fs)
    nfft = 1 << int(np.ceil(np.log2(n + L)))
    premix = np.empty((n_src, n_mics, n))
    for s in range(n_src):
        sig = _speechlike(rng, n, fs, env_shape, env_block)
        S = np.fft.rfft(sig, nfft)
        for m in range(n_mics):
            H = np.fft.rfft(_rir(rng, fs, rt60, L), nfft)
            premix[s, m] = np.fft.irfft(S * H, nfft)[:n]
    premix /= np.std(premix[:, 0, :], axis=1)[:, None, None]  # unit variance at the ref mic
    var_total = float(n_targets)
    sigma_n = np.sqrt(10 ** (-snr_db / 10) * var_total)
    sigma_i = np.sqrt(max(0.0, 10 ** (-sinr_db / 10) * var_total - sigma_n**2) / n_interferers)
    premix[n_targets:] *= sigma_i
    background = premix[n_targets:].sum(axis=0) + sigma_n * rng.standard_normal((n_mics, n))
    mix = premix[:n_targets].sum(axis=0) + background
    images = premix[:n_targets].transpose(0, 2, 1).copy()
    return mix.T.copy(), images


def stft_domain_mixture(seed, n_frames, n_freq, n_mics, n_src, noise_db=-60.0, n_interferers=10, sinr_db=10.0,
                        dtype=np.complex128):
    """X (T, F, M) drawn directly in the STFT domain: ``n_src`` complex-Laplacian targets plus
    ``n_interferers`` weaker sources (total SINR ``sinr_db``), each with a per-frame Gamma(0.5) activity
    shared across bins (the dependence the IVA source model exploits), a random complex mixing matrix
    per bin, and a white noise floor ``noise_db`` below the targets.  Same recipe as
    :func:`stft_domain_batch_torch`; well conditioned for the laplace model (a 1e-15 perturbation of X
    moves the reference's W by ~1e-14), NOT guaranteed for gauss -- parity tests use
    :func:`small_test_mixture` / :func:`convolutive_mixture` instead."""
    rng = np.random.default_rng(seed)
    T, F, M, K = n_frames, n_freq, n_mics, n_src
    Q = K + n_interferers
    act = rng.gamma(0.5, 1.0, size=(T, 1, Q)) + 0.05
    S = (rng.laplace(size=(T, F, Q)) + 1j * rng.laplace(size=(T, F, Q))) * act
    A = rng.standard_normal((F, M, Q)) + 1j * rng.standard_normal((F, M, Q))
    if n_interferers:
        A[:, :, K:] *= np.sqrt(10 ** (-sinr_db / 10) * K / n_interferers)
    X = np.einsum("fmk,tfk->tfm", A, S)
    sig = 10 ** (noise_db / 20) * np.sqrt(K)
    X += sig * (rng.standard_normal((T, F, M)) + 1j * rng.standard_normal((T, F, M)))
    return X.astype(dtype)


def small_test_mixture(seed, n_mics, n_targets, n_samples=2700, fs=8000, frame=64, hop=32,
                       n_interferers=4, rt60=0.02, env_shape=2.0, env_block=0.02,
                       dtype=np.complex128):
    """A small but *realistic* (time-domain convolutive, then STFT) input for parity tests:
    ``X (T, F, M)`` with ``T = (n_samples - frame)//hop + 1`` and ``F = frame//2 + 1``.  Real STFTs of
    convolutive mixtures keep the reference well conditioned for both source models (a 1e-15 relative
    perturbation of X moves W by ~1e-14 after 20 iterations), which direct STFT-domain draws do not
    guarantee for the gauss model."""
    mix, _ = convolutive_mixture(seed, n_mics, n_targets, duration=n_samples / fs, fs=fs,
                                 n_interferers=n_interferers, rt60=rt60,
                                 env_shape=env_shape, env_block=env_block)
    return stft(mix, frame, hop).astype(dtype)


def stft_domain_batch_torch(n_batch, n_frames, n_freq, n_mics, n_src, seed, device, dtype=None,
                            n_interferers=10, sinr_db=10.0, noise_db=-60.0, chunk=32):
    """Device-side generator for large throughput runs: (B, T, F, M) complex tensor drawn directly in the
    STFT domain -- per mixture ``n_src`` targets + ``n_interferers`` weaker sources, all complex Laplacian
    with a per-frame Gamma(0.5) activity shared across bins (the dependence the IVA source model exploits),
    a random complex mixing matrix per bin and a white noise floor.  Seeded per rank/step by ``seed``."""
    import torch

    dtype = dtype or torch.complex128
    rdt = torch.float64 if dtype == torch.complex128 else torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    B, T, F, M, K = n_batch, n_frames, n_freq, n_mics, n_src
    Q = K + n_interferers
    out = torch.empty((B, T, F, M), dtype=dtype, device=device)
    gain = torch.ones(Q, dtype=rdt, device=device)
    if n_interferers:
        gain[K:] = (10 ** (-sinr_db / 10) * K / n_interferers) ** 0.5
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)

        def rnd(*shape):
            return torch.randn(*shape, generator=g, device=device, dtype=rdt)

        act = 0.5 * rnd(nb, T, 1, Q) ** 2 + 0.05  # Gamma(0.5, 1) = chi^2_1 / 2
        lap = lambda *s: (torch.rand(*s, generator=g, device=device, dtype=rdt).clamp_min(1e-12).log()
                          - torch.rand(*s, generator=g, device=device, dtype=rdt).clamp_min(1e-12).log())
        S = torch.complex(lap(nb, T, F, Q), lap(nb, T, F, Q)) * (act * gain).to(dtype)
        A = torch.complex(rnd(nb, F, M, Q), rnd(nb, F, M, Q))
        X = torch.einsum("bfmq,btfq->btfm", A, S)
        sig = 10 ** (noise_db / 20) * K**0.5
        X += sig * torch.complex(rnd(nb, T, F, M), rnd(nb, T, F, M))
        out[b0 : b0 + nb] = X
    return out


def stft_domain_bins_torch(n_frames, n_freq_total, f_begin, f_end, n_mics, n_src, seed, device, dtype=None,
                           n_interferers=10, sinr_db=10.0, noise_db=-60.0):
    """Bins ``[f_begin, f_end)`` of ONE long synthetic mixture ``(T, f_end - f_begin, M)``, reproducible for ANY split
    of the frequency axis on multiples of 32 bins: the per-frame source activities depend on ``seed`` only and every
    group of 32 bins draws from its own generator ``(seed, group)``, so concatenating the ranges drawn by different
    ranks (or drawing the full range in one piece) gives bit-identical data -- what the frequency-sharded benchmark
    needs to compare its shards with a single-GPU run.  Same source model as :func:`stft_domain_batch_torch`."""
    import torch

    assert f_begin % 32 == 0 and 0 <= f_begin <= f_end <= n_freq_total
    dtype = dtype or torch.complex128
    rdt = torch.float64 if dtype == torch.complex128 else torch.float32
    T, M, K = n_frames, n_mics, n_src
    Q = K + n_interferers
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    act = 0.5 * torch.randn(T, 1, Q, generator=g, device=device, dtype=rdt) ** 2 + 0.05
    gain = torch.ones(Q, dtype=rdt, device=device)
    if n_interferers:
        gain[K:] = (10 ** (-sinr_db / 10) * K / n_interferers) ** 0.5
    amp = (act * gain).to(dtype)
    out = torch.empty((T, f_end - f_begin, M), dtype=dtype, device=device)
    sig = 10 ** (noise_db / 20) * K**0.5
    for f0 in range(f_begin, f_end, 32):
        nb = min(32, f_end - f0, n_freq_total - f0)
        g.manual_seed(int(seed) * 1000003 + 7919 * (f0 // 32) + 1)

        def rnd(*shape):
            return torch.randn(*shape, generator=g, device=device, dtype=rdt)

        def lap(*shape):
            u = torch.rand(2, *shape, generator=g, device=device, dtype=rdt).clamp_min(1e-12).log()
            return u[0] - u[1]

        # always draw a full group of 32 bins (so a ragged last group sees the same numbers however it is reached)
        S = torch.complex(lap(T, 32, Q), lap(T, 32, Q)) * amp
        A = torch.complex(rnd(32, M, Q), rnd(32, M, Q))
        X = torch.einsum("fmq,tfq->tfm", A, S)
        X += sig * torch.complex(rnd(T, 32, M), rnd(T, 32, M))
        out[:, f0 - f_begin : f0 - f_begin + nb] = X[:, :nb]
    return out


def audio_batch_torch(n_batch, n_samples, n_mics, n_src, seed, device, n_interferers=10, sinr_db=10.0, snr_db=60.0,
                      fs=16000, chunk=64):
    """Device-side generator of time-domain inputs for throughput runs of the audio-in / audio-out path:
    (B, N, M) float64 -- per mixture ``n_src`` targets and ``n_interferers`` weaker sources (super-Gaussian samples
    with a 250 ms block envelope), each reaching the microphones through a short random decaying filter (32 taps,
    applied as a sum of delayed copies), plus sensor noise; SINR / SNR as in ``convolutive_mixture``.  Cheap on
    purpose: the cost of the separation does not depend on the signal content."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    B, N, M, Q = n_batch, n_samples, n_mics, n_src + n_interferers
    out = torch.empty((B, N, M), dtype=torch.float64, device=device)
    blk = max(1, int(0.25 * fs))
    gain = torch.ones(Q, dtype=torch.float64, device=device)
    sig_n = (10 ** (-snr_db / 10) * n_src) ** 0.5
    if n_interferers:
        gain[n_src:] = (max(0.0, 10 ** (-sinr_db / 10) * n_src - sig_n**2) / n_interferers) ** 0.5
    taps = 32
    decay = torch.exp(-0.25 * torch.arange(taps, dtype=torch.float64, device=device))
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        rnd = lambda *s: torch.randn(*s, generator=g, device=device, dtype=torch.float64)
        S = rnd(nb, N, Q)
        S = S * S.abs()  # heavier tails than a Gaussian
        env = (0.5 * rnd(nb, N // blk + 1, Q) ** 2 + 0.05).repeat_interleave(blk, dim=1)[:, :N]
        S = S * env * gain
        H = rnd(nb, taps, Q, M) * decay[None, :, None, None] * 0.3
        H[:, 0] += 1.0
        x = sig_n * rnd(nb, N, M)
        for d in range(taps):
            x[:, d:] += torch.bmm(S[:, : N - d], H[:, d])
        out[b0 : b0 + nb] = x
    return out
