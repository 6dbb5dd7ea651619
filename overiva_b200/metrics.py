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
