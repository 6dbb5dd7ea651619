"""SDR / SIR of separated signals against known source images.

``mir_eval.separation.bss_eval_sources`` (what the reference's drivers call:
``overiva_oneshot.py:263-284,391-397``, ``overiva_sim.py:210-232``) is not available offline, so this
module implements the same decomposition with a one-tap (instantaneous) distortion filter:
``s_hat = s_target + e_interf + e_artif`` with ``s_target`` the projection of the estimate on its
own reference and ``s_target + e_interf`` the projection on the span of all references; the
output-to-reference permutation is resolved by maximising the mean SIR, as bss_eval does.  The SAME
routine scores both the CUDA path and the oracle, so "SDR/SIR within 0.1 dB" is well defined.
"""
from __future__ import annotations

import itertools

import numpy as np


def _decompose(est, refs, k):
    """est (N,), refs (K, N): returns (sdr_db, sir_db) of ``est`` scored as source k."""
    G = refs @ refs.T
    proj_all = np.linalg.solve(G, refs @ est) @ refs
    s_target = (refs[k] @ est) / G[k, k] * refs[k]
    e_interf = proj_all - s_target
    e_artif = est - proj_all
    tiny = np.finfo(float).tiny
    p_t = float(s_target @ s_target)
    sdr = 10 * np.log10(p_t / max(float((e_interf + e_artif) @ (e_interf + e_artif)), tiny))
    sir = 10 * np.log10(p_t / max(float(e_interf @ e_interf), tiny))
    return sdr, sir


def bss_eval(refs, ests):
    """refs (K, N) reference images at the reference mic, ests (J, N) separated signals (J >= K is
    allowed: the best K outputs are matched).  Returns ``(sdr (K,), sir (K,), perm (K,))`` where
    ``ests[perm[k]]`` is the estimate of source k."""
    refs = np.asarray(refs, dtype=np.float64)
    ests = np.asarray(ests, dtype=np.float64)
    n = min(refs.shape[1], ests.shape[1])
    refs, ests = refs[:, :n], ests[:, :n]
    K, J = refs.shape[0], ests.shape[0]
    sdr = np.empty((J, K))
    sir = np.empty((J, K))
    for j in range(J):
        for k in range(K):
            sdr[j, k], sir[j, k] = _decompose(ests[j], refs, k)
    best, best_perm = -np.inf, None
    for perm in itertools.permutations(range(J), K):
        score = np.mean([sir[perm[k], k] for k in range(K)])
        if score > best:
            best, best_perm = score, perm
    perm = np.array(best_perm)
    ks = np.arange(K)
    return sdr[perm, ks], sir[perm, ks], perm


# ---------------------------------------------------------------------------------------------------------------
# BSS Eval v3 with time-invariant distortion FILTERS -- what ``mir_eval.separation.bss_eval_sources`` computes (the
# metric of the reference's drivers and of its published ``data.json``; mir_eval is third-party and absent offline,
# so this restates the published algorithm of Vincent, Gribonval and Fevotte, "Performance measurement in blind audio
# source separation", IEEE TASLP 2006: PARITY UNPINNED against mir_eval itself).  ``bss_eval`` above is its
# ``flen = 1`` special case and is what the device-side monitor evaluates; this one is for final scoring on the host.
# ---------------------------------------------------------------------------------------------------------------
def _delayed_gram(refs, flen):
    """G[(k, a), (l, b)] = sum_n refs[k][n - a] refs[l][n - b] for 0 <= a, b < flen (block Toeplitz), via FFT."""
    K, N = refs.shape
    nfft = 1 << int(np.ceil(np.log2(N + flen)))
    Rf = np.fft.rfft(refs, nfft, axis=1)
    G = np.empty((K * flen, K * flen))
    idx = np.arange(flen)
    lag = idx[:, None] - idx[None, :]  # a - b
    for k in range(K):
        for l in range(k, K):
            # c[m] = sum_n refs[k][n] refs[l][n + m]  =>  entry (a, b) = c[a - b]
            c = np.fft.irfft(np.conj(Rf[k]) * Rf[l], nfft)
            blk = c[lag % nfft]
            G[k * flen : (k + 1) * flen, l * flen : (l + 1) * flen] = blk
            if l != k:
                G[l * flen : (l + 1) * flen, k * flen : (k + 1) * flen] = blk.T
    return G


def _project(refs, est, flen, G=None):
    """Least-squares projection of ``est`` on the span of the references delayed by 0..flen-1 samples; signals of
    length N are treated as zero-padded to N + flen - 1 (as mir_eval does).  Returns the projection (N + flen - 1,)."""
    K, N = refs.shape
    nfft = 1 << int(np.ceil(np.log2(N + flen)))
    if G is None:
        G = _delayed_gram(refs, flen)
    Rf = np.fft.rfft(refs, nfft, axis=1)
    Ef = np.fft.rfft(est, nfft)
    D = np.empty(K * flen)
    for k in range(K):
        c = np.fft.irfft(np.conj(Rf[k]) * Ef, nfft)  # c[a] = sum_n refs[k][n] est[n + a]
        D[k * flen : (k + 1) * flen] = c[:flen]
    try:
        C = np.linalg.solve(G, D)
    except np.linalg.LinAlgError:
        C = np.linalg.lstsq(G, D, rcond=None)[0]
    C = C.reshape(K, flen)
    out = np.zeros(N + flen - 1)
    for k in range(K):
        out += np.convolve(refs[k], C[k])[: N + flen - 1]
    return out


def bss_eval_sources(refs, ests, flen=512, compute_permutation=True):
    """``mir_eval.separation.bss_eval_sources(reference_sources, estimated_sources)`` restated: refs (K, N), ests
    (K, N) -> ``(sdr, sir, sar, perm)`` with ``ests[perm[k]]`` the estimate of source k (the permutation maximising
    the mean SIR, as mir_eval does).  Distortion filters of ``flen`` taps (mir_eval's default: 512)."""
    refs = np.atleast_2d(np.asarray(refs, dtype=np.float64))
    ests = np.atleast_2d(np.asarray(ests, dtype=np.float64))
    if refs.shape != ests.shape:
        raise ValueError("refs and ests must have the same shape, got %s and %s" % (refs.shape, ests.shape))
    K, N = refs.shape
    tiny = np.finfo(float).tiny
    G = _delayed_gram(refs, flen)
    sdr = np.empty((K, K))
    sir = np.empty((K, K))
    sar = np.empty((K, K))
    pad = np.zeros(flen - 1)
    for j in range(K):  # estimate
        e = np.concatenate([ests[j], pad])
        p_all = _project(refs, ests[j], flen, G)
        for k in range(K):  # scored as source k
            s_t = _project(refs[k : k + 1], ests[j], flen, G[k * flen : (k + 1) * flen, k * flen : (k + 1) * flen])
            e_i = p_all - s_t
            e_a = e - p_all
            pt = float(s_t @ s_t)
            sdr[j, k] = 10 * np.log10(max(pt, tiny) / max(float((e_i + e_a) @ (e_i + e_a)), tiny))
            sir[j, k] = 10 * np.log10(max(pt, tiny) / max(float(e_i @ e_i), tiny))
            sar[j, k] = 10 * np.log10(max(float((s_t + e_i) @ (s_t + e_i)), tiny) / max(float(e_a @ e_a), tiny))
    if compute_permutation:
        best, perm = -np.inf, None
        for cand in itertools.permutations(range(K)):
            score = np.mean([sir[cand[k], k] for k in range(K)])
            if score > best:
                best, perm = score, cand
        perm = np.array(perm)
    else:
        perm = np.arange(K)
    ks = np.arange(K)
    return sdr[perm, ks], sir[perm, ks], sar[perm, ks], perm


def xcorr_reference(refs, ests, flen):
    """numpy statement of what ``oiva_xcorr`` computes: c (K, K + J, flen) with ``c[i, j, m] = sum_n s_i[n] s_j[n + m]``
    (s_i a reference, s_j a reference or an estimate), and the energies ee (J,) of the estimates."""
    refs = np.atleast_2d(np.asarray(refs, dtype=np.float64))
    ests = np.atleast_2d(np.asarray(ests, dtype=np.float64))
    K, N = refs.shape
    allsig = np.concatenate([refs, ests], axis=0)
    nfft = 1 << int(np.ceil(np.log2(N + flen)))
    Af = np.fft.rfft(allsig, nfft, axis=1)
    c = np.empty((K, allsig.shape[0], flen))
    for i in range(K):
        c[i] = np.fft.irfft(np.conj(Af[i])[None] * Af, nfft, axis=1)[:, :flen]
    return c, np.einsum("jn,jn->j", ests, ests)


def bss_eval_sources_from_xcorr(c, ee, flen=None, compute_permutation=True):
    """BSS Eval v3 with ``flen``-tap distortion filters from cross-correlations only.  c: (K, K + J, flen) as
    ``xcorr_reference`` / the device kernel ``oiva_xcorr`` return it, ee: (J,) energies of the estimates.  Returns
    ``(sdr, sir, sar, perm)`` exactly as ``bss_eval_sources``: with D the references delayed by 0..flen-1 samples,
    G = D D^T is block Toeplitz in c[k, l], b_j = D e_j is c[k, K + j], and
    ``|s_target|^2 = b_jk^T G_kk^-1 b_jk``, ``|P e_j|^2 = b_j^T G^-1 b_j``, ``|e_interf|^2 = |P e|^2 - |s_target|^2``,
    ``|e_artif|^2 = |e|^2 - |P e|^2``, ``|e_interf + e_artif|^2 = |e|^2 - |s_target|^2`` (s_target and P e are
    orthogonal projections of e on nested subspaces)."""
    c = np.asarray(c, dtype=np.float64)
    ee = np.asarray(ee, dtype=np.float64)
    K, R, fl = c.shape
    flen = fl if flen is None else int(flen)
    J = R - K
    idx = np.arange(flen)
    lag = idx[:, None] - idx[None, :]
    pos, alag = lag >= 0, np.abs(lag)
    G = np.empty((K * flen, K * flen))
    for k in range(K):
        for l in range(K):
            G[k * flen : (k + 1) * flen, l * flen : (l + 1) * flen] = np.where(pos, c[k, l][alag], c[l, k][alag])
    B = c[:, K:, :flen].transpose(0, 2, 1).reshape(K * flen, J)  # column j: b_j

    def quad(Gm, Bm):  # b^T G^-1 b per column
        try:
            X = np.linalg.solve(Gm, Bm)
        except np.linalg.LinAlgError:
            X = np.linalg.lstsq(Gm, Bm, rcond=None)[0]
        return np.einsum("ij,ij->j", Bm, X)

    pe2 = quad(G, B)
    tiny = np.finfo(float).tiny
    sdr = np.empty((J, K))
    sir = np.empty((J, K))
    sar = np.empty((J, K))
    for k in range(K):
        sl = slice(k * flen, (k + 1) * flen)
        st2 = quad(G[sl, sl], B[sl])
        for j in range(J):
            sdr[j, k] = 10 * np.log10(max(st2[j], tiny) / max(ee[j] - st2[j], tiny))
            sir[j, k] = 10 * np.log10(max(st2[j], tiny) / max(pe2[j] - st2[j], tiny))
            sar[j, k] = 10 * np.log10(max(pe2[j], tiny) / max(ee[j] - pe2[j], tiny))
    if compute_permutation:
        best, perm = -np.inf, None
        for cand in itertools.permutations(range(J), K):
            score = np.mean([sir[cand[k], k] for k in range(K)])
            if score > best:
                best, perm = score, cand
        perm = np.array(perm)
    else:
        perm = np.arange(K)
    ks = np.arange(K)
    return sdr[perm, ks], sir[perm, ks], sar[perm, ks], perm
