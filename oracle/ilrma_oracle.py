"""CPU oracle for ILRMA -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

**Parity unpinned.**  The reference calls ``pyroomacoustics.bss.ilrma`` (overiva_oneshot.py:331-339,
overiva_sim.py:309-311); pyroomacoustics (pinned ==0.1.23, environment.yml:13) is third-party and absent from
/root/reference, and there is no network to fetch it.  This file restates the published algorithm (D. Kitamura, N. Ono,
H. Sawada, H. Kameoka, H. Saruwatari, "Determined blind source separation unifying independent vector analysis and
nonnegative matrix factorization", IEEE/ACM TASLP 2016) in the form pyroomacoustics implements it: multiplicative NMF
updates of a rank-``n_components`` spectrogram model per source, iterative-projection updates of the demixing vectors
with the covariance weighted by 1 / r_k(f, t), a scale normalisation by the mean output power after every epoch.
Conventions of this repository: ``W`` is (n_freq, n_chan, n_src) with the demixing vectors in the COLUMNS
(y_k = w_k^H x, as overiva.py:135-136), the callback cadence is overiva's (every 10th epoch, before the update), the random
initial NMF factors are drawn from numpy's global generator in the order T then V (pass T0 / V0 to fix them).
"""
import numpy as np

from .overiva_oracle import demix, projection_back


def ilrma(X, n_src=None, n_iter=20, proj_back=False, W0=None, n_components=2, return_filters=False, callback=None,
          T0=None, V0=None):
    n_frames, n_freq, n_chan = X.shape
    if n_src is None:
        n_src = n_chan
    if n_src != n_chan:
        raise ValueError("ILRMA is a determined algorithm: n_src must equal the number of channels")
    Xc = X.astype(np.complex128)
    W = np.zeros((n_freq, n_chan, n_src), dtype=np.complex128)
    if W0 is None:
        W[:, :, :] = np.eye(n_chan, n_src)
    else:
        W[:, :, :] = W0
    T = np.array(T0, dtype=np.float64) if T0 is not None else 0.1 + 0.9 * np.random.rand(n_src, n_freq, n_components)
    V = np.array(V0, dtype=np.float64) if V0 is not None else 0.1 + 0.9 * np.random.rand(n_src, n_frames, n_components)
    eps = 1e-15
    Xf = np.ascontiguousarray(Xc.swapaxes(0, 1))  # (F, T, M)
    R = np.matmul(T, V.swapaxes(1, 2))  # (K, F, T)
    Y = demix(Xf, W)  # (F, T, K)
    P = np.abs(Y.transpose(2, 0, 1)) ** 2  # (K, F, T)
    eye = np.eye(n_chan)
    for epoch in range(n_iter):
        if callback is not None and epoch % 10 == 0:
            Yt = Y.swapaxes(0, 1)
            if proj_back:
                z = projection_back(Yt, Xc[:, :, 0])
                callback((Yt * np.conj(z[None])).astype(X.dtype))
            else:
                callback(Yt.astype(X.dtype))
        for s in range(n_src):
            iR = 1.0 / R[s]
            T[s] *= np.sqrt(np.dot(P[s] * iR ** 2, V[s]) / np.dot(iR, V[s]))
            T[s][T[s] < eps] = eps
            R[s] = np.dot(T[s], V[s].T)
            iR = 1.0 / R[s]
            V[s] *= np.sqrt(np.dot((P[s] * iR ** 2).T, T[s]) / np.dot(iR.T, T[s]))
            V[s][V[s] < eps] = eps
            R[s] = np.dot(T[s], V[s].T)
            iR = 1.0 / R[s]
            # auxiliary variable C[f] = (1/T) sum_t x x^H / r_s(f, t), then the IP update of w_s
            C = np.matmul((Xf * iR[:, :, None]).swapaxes(1, 2), np.conj(Xf)) / n_frames  # (F, M, M)
            WV = np.matmul(np.conj(W.swapaxes(1, 2)), C)  # rows w_k^H C
            w = np.linalg.solve(WV, np.broadcast_to(eye[:, s, None], (n_freq, n_chan, 1)))[:, :, 0]
            denom = np.einsum("fm,fmn,fn->f", np.conj(w), C, w)
            W[:, :, s] = w / np.sqrt(denom)[:, None]
        Y = demix(Xf, W)
        P = np.abs(Y.transpose(2, 0, 1)) ** 2
        for s in range(n_src):
            lam = 1.0 / np.sqrt(np.mean(P[s]))
            W[:, :, s] *= lam
            P[s] *= lam ** 2
            R[s] *= lam ** 2
            T[s] *= lam ** 2
    Yt = np.ascontiguousarray(Y.swapaxes(0, 1))
    if proj_back:
        z = projection_back(Yt, Xc[:, :, 0])
        Yt = Yt * np.conj(z[None])
    Yt = Yt.astype(X.dtype)
    if return_filters:
        return Yt, W
    return Yt
