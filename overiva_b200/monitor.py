"""Device-side SDR / SIR convergence monitor for ``callback=``.

The reference's drivers monitor convergence with a callback that runs every 10 iterations on the CPU: inverse STFT of
the current estimate, reorder by power, drop the ``framesize // 2`` delay, ``mir_eval.separation.bss_eval_sources``
against the clean source images (``overiva_oneshot.py:263-284``, ``overiva_sim.py:210-232``).  With the loop on the
GPU that would mean a device-to-host copy of the whole estimate every 10 epochs.  :class:`ConvergenceMonitor` is the
same callback kept on the device: the estimate arrives as a CUDA tensor (entry points hand CUDA tensors to the
callback when X is a CUDA tensor), the inverse STFT is ``overiva_b200.stft.synthesis``, and the only pass over the
audio that the metric needs -- the Gram matrix of references and estimates -- is one small kernel (``oiva_gram``);
a (2K+2) x (2K+2) matrix goes to the host.

The metric is ``overiva_b200.metrics.bss_eval`` (mir_eval is not available offline: same decomposition with a
one-tap distortion filter), evaluated from the Gram matrix: with G = R R^T, b = R e, |e|^2,
``|s_target|^2 = b_k^2 / G_kk``, ``|P e|^2 = b^T G^-1 b``, ``SDR = |s_t|^2 / (|e|^2 - |s_t|^2)``,
``SIR = |s_t|^2 / (|P e|^2 - |s_t|^2)``.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from . import _lib as L
from . import core, stft


def gram(a, b=None):
    """Gram matrix [a; b][a; b]^T of real float64 CUDA signals a (Ra, N), b (Rb, N) (any strides) -> (R, R) CUDA."""
    lib = L.load()
    if not (a.is_cuda and a.dtype == torch.float64 and a.ndim == 2):
        raise TypeError("gram needs 2-D float64 CUDA tensors")
    if b is not None and not (b.is_cuda and b.dtype == torch.float64 and b.ndim == 2 and b.shape[1] == a.shape[1]):
        raise TypeError("gram needs 2-D float64 CUDA tensors of equal length")
    Ra, N = a.shape
    Rb = 0 if b is None else b.shape[0]
    dev = a.device
    with torch.cuda.device(dev):
        scratch = torch.empty(lib.oiva_gram_scratch_bytes(Ra + Rb, N), dtype=torch.uint8, device=dev)
        out = torch.empty((Ra + Rb, Ra + Rb), dtype=torch.float64, device=dev)
        L.check(lib.oiva_gram(core._ptr(a), a.stride(0), a.stride(1), Ra, core._ptr(b),
                              0 if b is None else b.stride(0), 0 if b is None else b.stride(1), Rb, N,
                              core._ptr(scratch), core._ptr(out), core._stream_ptr(dev)), "oiva_gram")
    return out


def bss_eval_from_gram(G, K):
    """G: (K + J, K + J) Gram matrix of K references followed by J estimates -> (sdr (K,), sir (K,), perm (K,)),
    the same quantities as ``metrics.bss_eval(refs, ests)``."""
    G = np.asarray(G, dtype=np.float64)
    J = G.shape[0] - K
    Grr, Gre = G[:K, :K], G[:K, K:]
    ee = np.diag(G)[K:]
    proj = np.einsum("kj,kj->j", Gre, np.linalg.solve(Grr, Gre))  # |P e_j|^2
    tiny = np.finfo(float).tiny
    sdr = np.empty((J, K))
    sir = np.empty((J, K))
    for j in range(J):
        for k in range(K):
            p_t = Gre[k, j] ** 2 / Grr[k, k]
            sdr[j, k] = 10 * np.log10(p_t / max(ee[j] - p_t, tiny)) if p_t > 0 else -np.inf
            sir[j, k] = 10 * np.log10(p_t / max(proj[j] - p_t, tiny)) if p_t > 0 else -np.inf
    best, best_perm = -np.inf, None
    for perm in itertools.permutations(range(J), K):
        score = np.mean([sir[perm[k], k] for k in range(K)])
        if score > best:
            best, best_perm = score, perm
    perm = np.array(best_perm)
    ks = np.arange(K)
    return sdr[perm, ks], sir[perm, ks], perm


def bss_eval_device(refs, ests):
    """``metrics.bss_eval`` for CUDA tensors refs (K, N), ests (J, N) (float64, any strides; the shorter length is
    used): one Gram kernel on the device, (K+J)^2 doubles to the host."""
    n = min(refs.shape[1], ests.shape[1])
    G = gram(refs[:, :n], ests[:, :n]).cpu().numpy()
    return bss_eval_from_gram(G, refs.shape[0])


def xcorr(refs, ests, flen):
    """Cross-correlations at lags 0..flen-1 on the device (``oiva_xcorr``): refs (K, N), ests (J, N) float64 CUDA tensors
    (any strides) -> c (K, K + J, flen) CUDA with ``c[i, j, m] = sum_n s_i[n] s_j[n + m]``."""
    lib = L.load()
    for t in (refs, ests):
        if not (t.is_cuda and t.dtype == torch.float64 and t.ndim == 2):
            raise TypeError("xcorr needs 2-D float64 CUDA tensors")
    if refs.shape[1] != ests.shape[1]:
        raise ValueError("refs and ests must have the same length")
    K, N = refs.shape
    J = ests.shape[0]
    dev = refs.device
    with torch.cuda.device(dev):
        scratch = torch.empty(lib.oiva_xcorr_scratch_bytes(K, J, N, int(flen)), dtype=torch.uint8, device=dev)
        out = torch.empty((K, K + J, int(flen)), dtype=torch.float64, device=dev)
        L.check(lib.oiva_xcorr(core._ptr(refs), refs.stride(0), refs.stride(1), K, core._ptr(ests), ests.stride(0),
                               ests.stride(1), J, N, int(flen), core._ptr(scratch), core._ptr(out), core._stream_ptr(dev)),
                "oiva_xcorr")
    return out


def bss_eval_sources_device(refs, ests, flen=512):
    """``mir_eval.separation.bss_eval_sources`` (restated in ``metrics.bss_eval_sources``: distortion filters of ``flen``
    taps) for CUDA tensors refs (K, N), ests (J, N): the pass over the audio -- cross-correlations at ``flen`` lags --
    runs on the device, ``(K + J) K flen`` doubles go to the host, where the block-Toeplitz systems are solved.
    -> (sdr, sir, sar, perm)."""
    from . import metrics

    n = min(refs.shape[1], ests.shape[1])
    r, e = refs[:, :n], ests[:, :n]
    c = xcorr(r, e, flen).cpu().numpy()
    ee = torch.einsum("jn,jn->j", e, e).cpu().numpy()
    return metrics.bss_eval_sources_from_xcorr(c, ee, flen)


class ConvergenceMonitor:
    """``callback`` object for ``overiva`` / ``auxiva_pca`` / ``ogive`` (pass X as a CUDA tensor so that the
    estimates stay on the device).  After the run ``SDR`` / ``SIR`` hold one array per call, like the lists the
    reference's ``convergence_callback`` fills (``overiva_oneshot.py:261-284``).

    ref: (n_src_ref, n_samples, n_mics) or (n_src_ref, n_samples) clean source images (CUDA or numpy); scored at
    microphone 0.  ``delay``: samples to drop at the head of the synthesised estimate (``framesize // 2`` when the
    analysis used the ``L - hop`` state-buffer padding, as the reference's STFT does; 0 for un-padded framing).
    ``filter_length``: taps of the metric's distortion filters (1, or 512 for mir_eval's ``bss_eval_sources``)."""

    def __init__(self, ref, framesize=4096, hop=None, win_s=None, n_targets=None, reorder=True, delay=0, device=None,
                 filter_length=1):
        dev = core._require_cuda(device)
        r = ref if isinstance(ref, torch.Tensor) else torch.from_numpy(np.asarray(ref, dtype=np.float64))
        if r.ndim == 3:
            r = r[:, :, 0]
        self.ref = r.to(device=dev, dtype=torch.float64).contiguous()
        self.L = int(framesize)
        self.hop = self.L // 2 if hop is None else int(hop)
        self.win_s = stft.compute_synthesis_window(stft.hann(self.L), self.hop) if win_s is None else win_s
        self.n_targets = self.ref.shape[0] if n_targets is None else int(n_targets)
        self.reorder = bool(reorder)
        self.delay = int(delay)
        # distortion-filter taps of the metric: 1 = the instantaneous decomposition (metrics.bss_eval, a Gram matrix);
        # 512 = what the reference's callback computes with mir_eval (cross-correlations at 512 lags on the device)
        self.filter_length = int(filter_length)
        self.SDR, self.SIR = [], []

    def __call__(self, Y, **kwargs):
        Yt = Y if isinstance(Y, torch.Tensor) else torch.from_numpy(np.asarray(Y))
        Yt = Yt.to(self.ref.device)
        y = stft.synthesis(Yt, self.L, self.hop, win=self.win_s).to(torch.float64)  # (N', K)
        if self.reorder:  # decreasing power, overiva_oneshot.py:276-278
            y = y[:, torch.argsort(y.std(dim=0), descending=True)]
        m = min(y.shape[0] - self.delay, self.ref.shape[1])
        k = self.n_targets
        if self.filter_length > 1:
            sdr, sir, _, _ = bss_eval_sources_device(self.ref[:k, :m], y[self.delay : self.delay + m, :k].T, self.filter_length)
        else:
            sdr, sir, _ = bss_eval_device(self.ref[:k, :m], y[self.delay : self.delay + m, :k].T)
        self.SDR.append(sdr)
        self.SIR.append(sir)
