"""CPU oracle for the OverIVA demixing loop -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a numpy restatement of the algorithm implemented by the reference
repository onolab-tmu/overiva (``overiva.py``, ``auxiva_pca.py``, ``ive.py``) plus the one
third-party routine on the path, ``pyroomacoustics.bss.projection_back``
(pyroomacoustics==0.1.23, pinned in the reference's ``environment.yml:13``; not vendored in
``/root/reference``, restated here from its published formula and from how the reference's
call sites use its result, ``overiva.py:198-199``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module, and only as the checker or as the timed CPU arm.
The product (``overiva_b200``) never imports it and has no CPU fallback.

Pinning status
--------------
The reference ships no tests, golden vectors or fixtures for this path.
The oracle is therefore pinned against *outputs of the reference itself run in the build
container*: ``oracle/validate_against_reference.py`` imports the unmodified
``/root/reference/{overiva,auxiva_pca,ive}.py`` under a two-part compatibility shim (a stub
``pyroomacoustics.bss.projection_back`` with the formula below; the numpy-1.x "stack of
vectors" rule for ``numpy.linalg.solve`` that ``overiva.py:182`` relies on) and asserts that
every function here agrees with it to <= 1e-12; ``tests/golden/make_golden.py`` stores the
*reference's* outputs as fixtures, and ``tests/test_oracle.py`` re-checks the oracle against
them wherever it runs.  ``projection_back`` itself remains "parity unpinned" (no copy of
pyroomacoustics is available offline); it is *defined* by the formula in its docstring.

The functions are split into the same steps the CUDA kernels implement, so that kernel-level
parity tests can compare intermediate quantities (covariances, source-model weights, one
iterative-projection sweep) and not only the end result.
"""
from __future__ import annotations

import numpy as np

EPS_R = 1e-15  # clamp on the source-model statistic, overiva.py:169-171 / ive.py:212-213


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def projection_back(Y, ref):
    """Least-squares scale of every separated channel onto a reference microphone.

    Restates ``pyroomacoustics.bss.projection_back(Y, ref)`` as used at
    ``overiva.py:145,198``, ``auxiva_pca.py:89`` and ``ive.py:197,250``:
    ``z[f,k] = sum_t conj(ref[t,f]) Y[t,f,k] / sum_t |Y[t,f,k]|^2`` and ``z = 1`` where the
    denominator is zero.  The callers then apply ``Y *= conj(z)``.

    Y: (T, F, K) complex, ref: (T, F) complex -> z: (F, K) complex.
    """
    num = np.einsum("tf,tfk->fk", np.conj(ref), Y)
    den = np.einsum("tfk,tfk->fk", np.conj(Y), Y).real
    z = np.ones(num.shape, dtype=np.result_type(Y.dtype, np.complex64))
    nz = den > 0.0
    z[nz] = num[nz] / den[nz]
    return z


def input_covariance(X):
    """``Cx[f] = (1/T) sum_t x(f,t) x(f,t)^H`` -- overiva.py:87, ive.py:97, auxiva_pca.py:71.

    X: (T, F, M) -> (F, M, M).  Evaluated per bin as a matrix product (identical maths, no
    (T, F, M, M) temporary, which is 118 GB at BASELINE config 5).
    """
    T = X.shape[0]
    Xf = X.transpose(1, 2, 0)  # (F, M, T)
    return (Xf @ np.conj(Xf.transpose(0, 2, 1))) / T


def herm(A):
    """Conjugate transpose of a stack of matrices (``tensor_H`` at overiva.py:93-94)."""
    return np.conj(np.swapaxes(A, -1, -2))


def solve_vec(A, b):
    """Stack-of-vectors solve: what ``np.linalg.solve(A, b)`` meant at overiva.py:182 under
    numpy 1.x (``b.ndim == A.ndim - 1``); numpy >= 2 needs the explicit column axis."""
    return np.linalg.solve(A, b[..., None])[..., 0]


def background_update(W_hat, Cx, n_src):
    """OverIVA's orthogonal-constraint refresh of the background block J.

    overiva.py:96-98: ``tmp = W^H Cx`` (K x M); ``J = tmp[:, :K]^{-1} tmp[:, K:]`` written
    into ``W_hat[:, :K, K:]``.  In place.
    """
    W = W_hat[:, :, :n_src]
    tmp = herm(W) @ Cx
    W_hat[:, :n_src, n_src:] = np.linalg.solve(tmp[:, :, :n_src], tmp[:, :, n_src:])


def principal_eig_init(Cx, n_src):
    """``init_eig`` branch, overiva.py:103-109: general ``eig`` of the (Hermitian) covariance,
    per bin keep the ``n_src`` columns with the largest eigenvalues in ascending order
    (``argsort(v)[-K:]``) and store their conjugate."""
    vals, vecs = np.linalg.eig(Cx)
    F, M, _ = Cx.shape
    W = np.empty((F, M, n_src), dtype=Cx.dtype)
    for f in range(F):
        keep = np.argsort(vals[f])[-n_src:]
        W[f] = np.conj(vecs[f][:, keep])
    return W


def init_demixing(Cx, n_src, W0=None, init_eig=False):
    """Build ``W_hat`` (F, M, M) exactly as overiva.py:89-123 does.

    Columns ``:K`` hold the demixing vectors w_k (``W``), ``W_hat[:, :K, K:] = J`` and
    ``W_hat[:, K:, K:] = -I`` so that ``W_hat^H = [W^H ; (J^H, -I)]``.
    """
    F, M, _ = Cx.shape
    W_hat = np.zeros((F, M, M), dtype=Cx.dtype)
    if W0 is not None:
        W_hat[:, :, :n_src] = W0  # overiva.py:116-117 (broadcast assignment into (F, M, K))
    elif init_eig:
        W_hat[:, :, :n_src] = principal_eig_init(Cx, n_src)
    else:
        idx = np.arange(n_src)
        W_hat[:, idx, idx] = 1.0  # overiva.py:111-114
    if n_src < M:
        background_update(W_hat, Cx, n_src)  # overiva.py:120-121
        idx = np.arange(n_src, M)
        W_hat[:, idx, idx] = -1.0  # overiva.py:122-123
    return W_hat


def demix(Xf, W):
    """``y_k(f,t) = w_k(f)^H x(f,t)``: overiva.py:135-136.  Xf (F,T,M), W (F,M,K) -> (F,T,K)."""
    return Xf @ np.conj(W)


def demix_power(Xf, W):
    """``r2[t,k] = sum_f |y_k(f,t)|^2`` -- the only cross-bin quantity of the loop
    (the argument of ``np.linalg.norm(Y, axis=0)`` at overiva.py:153-155)."""
    Y = demix(Xf, W)
    return np.einsum("ftk,ftk->tk", np.conj(Y), Y).real


def source_model(r2, model, n_freq):
    """overiva.py:152-173 from the squared statistic ``r2[t,k]``.

    Returns ``(r_inv (T,K), w_scale (K,))`` where ``W`` must be divided by ``w_scale``
    (gamma for laplace, sqrt(gamma) for gauss).  An unknown model string leaves r at zero,
    exactly as the reference does (then clamped to 1e-15); the rescale of W is skipped.
    """
    if model == "laplace":
        r = 2.0 * np.sqrt(r2)
    elif model == "gauss":
        r = r2 / n_freq
    else:
        r = np.zeros_like(r2)
    gamma = r.mean(axis=0)
    r = r / gamma[None, :]
    if model == "laplace":
        w_scale = gamma
    elif model == "gauss":
        w_scale = np.sqrt(gamma)
    else:
        w_scale = np.ones_like(gamma)
    r = np.where(r < EPS_R, EPS_R, r)
    return 1.0 / r, w_scale


def weighted_covariance(Xf, r_inv_s):
    """``V_s[f] = (1/T) sum_t x x^H / r_s(t)`` -- overiva.py:179.  Xf (F,T,M), r_inv_s (T,)."""
    T = Xf.shape[1]
    Xt = Xf.swapaxes(1, 2)  # (F, M, T)
    return ((Xt * r_inv_s[None, None, :]) @ np.conj(Xf)) / T


def ip_update_source(W_hat, V, Cx, s, n_src):
    """One iterative-projection update of source ``s`` in place: overiva.py:181-190."""
    M = W_hat.shape[1]
    WV = herm(W_hat) @ V
    e_s = np.zeros((W_hat.shape[0], M), dtype=W_hat.dtype)
    e_s[:, s] = 1.0
    w = solve_vec(WV, e_s)
    denom = np.einsum("fi,fij,fj->f", np.conj(w), V, w)
    W_hat[:, :, s] = w / np.sqrt(denom)[:, None]
    if n_src < M:
        background_update(W_hat, Cx, n_src)


def iterate_once(Xf, W_hat, Cx, n_src, model, n_freq_total=None, r2=None):
    """One epoch of the loop body overiva.py:138-190 (without the callback), in place on W_hat.

    ``r2`` may be supplied (frequency-sharded operation: the caller all-reduces the partial
    sums over bins); ``n_freq_total`` is the F that the gauss model divides by.
    """
    F = Xf.shape[0]
    if n_freq_total is None:
        n_freq_total = F
    W = W_hat[:, :, :n_src]
    if r2 is None:
        r2 = demix_power(Xf, W)
    r_inv, w_scale = source_model(r2, model, n_freq_total)
    W /= w_scale[None, None, :]  # overiva.py:161-167 (the Y /= ... there is dead work)
    for s in range(n_src):
        V = weighted_covariance(Xf, r_inv[:, s])
        ip_update_source(W_hat, V, Cx, s, n_src)
    return r_inv


# --------------------------------------------------------------------------------------
# entry points (same signatures as the reference)
# --------------------------------------------------------------------------------------
def overiva(
    X,
    n_src=None,
    n_iter=20,
    proj_back=True,
    W0=None,
    model="laplace",
    init_eig=False,
    return_filters=False,
    callback=None,
):
    """Restatement of ``overiva()`` -- overiva.py:28-204.

    X: (T, F, M) complex.  Returns Y (T, F, K) [and W (F, M, K), a view of W_hat, if
    ``return_filters``].  dtype follows X (overiva.py:89,126,131).
    """
    T, F, M = X.shape
    if n_src is None:
        n_src = M  # determined AuxIVA, overiva.py:83-84

    Cx = input_covariance(X).astype(X.dtype, copy=False)
    W_hat = init_demixing(Cx, n_src, W0=W0, init_eig=init_eig)
    W = W_hat[:, :, :n_src]
    Xf = np.ascontiguousarray(X.swapaxes(0, 1))  # (F, T, M), overiva.py:132

    for epoch in range(n_iter):
        if callback is not None and epoch % 10 == 0:  # overiva.py:142-148
            Y_tmp = demix(Xf, W).swapaxes(0, 1)
            if proj_back:
                z = projection_back(Y_tmp, X[:, :, 0])
                callback(Y_tmp * np.conj(z[None, :, :]))
            else:
                callback(Y_tmp)
        iterate_once(Xf, W_hat, Cx, n_src, model)

    Y = np.ascontiguousarray(demix(Xf, W).swapaxes(0, 1))  # overiva.py:192-194
    if proj_back:
        z = projection_back(Y, X[:, :, 0])
        Y *= np.conj(z[None, :, :])
    if return_filters:
        return Y, W
    return Y


def auxiva(X, **kwargs):
    """Determined AuxIVA = ``overiva`` with ``n_src`` omitted (overiva_oneshot.py:301-309)."""
    kwargs.pop("n_src", None)
    return overiva(X, n_src=None, **kwargs)


def auxiva_pca(X, n_src=None, **kwargs):
    """Restatement of ``auxiva_pca()`` -- auxiva_pca.py:63-92.

    PCA to ``n_src`` channels (``eigh``, last K eigenvectors), determined AuxIVA on the reduced
    signal with ``proj_back=False``, then projection back onto the *original* microphone 0.
    ``proj_back`` must be present in kwargs (auxiva_pca.py:86 pops it unconditionally).
    """
    T, F, M = X.shape
    if n_src is None:
        n_src = M
    if n_src < M:
        Cx = input_covariance(X)
        _, vecs = np.linalg.eigh(Cx)
        new_X = (X.swapaxes(0, 1) @ np.conj(vecs[:, :, -n_src:])).swapaxes(0, 1)
    else:
        new_X = X
    kwargs.pop("proj_back")
    Y = overiva(new_X, proj_back=False, **kwargs)
    z = projection_back(Y, X[:, :, 0])
    Y *= np.conj(z[None, :, :])
    return Y


def ogive(
    X,
    n_iter=4000,
    step_size=0.1,
    tol=1e-3,
    update="demix",
    proj_back=True,
    W0=None,
    model="laplace",
    init_eig=False,
    return_filters=False,
    callback=None,
):
    """Restatement of ``ogive()`` -- ive.py:33-256 (orthogonally constrained gradient IVE, K=1).

    The per-iteration statistic is written through the weighted covariance:
    ``x_psi = (sum_t x conj(y)/r) / (sum_t |y|^2/r) = V w / (w^H V w)`` with
    ``V = (1/T) sum_t x x^H / r(t)`` (ive.py:216-222) -- the same quantity the CUDA path uses.
    """
    T, F, M = X.shape
    Cx = input_covariance(X).astype(X.dtype, copy=False)  # ive.py:97
    Cx_inv = np.linalg.inv(Cx)  # ive.py:98
    Cx_norm = np.linalg.norm(Cx, axis=(1, 2))  # ive.py:99

    w = np.zeros((F, M, 1), dtype=X.dtype)
    a = np.zeros((F, M, 1), dtype=X.dtype)
    delta = np.zeros((F, M, 1), dtype=X.dtype)
    lambda_a = np.zeros((F, 1, 1), dtype=np.float64)

    if W0 is not None:  # ive.py:129-130
        w[:, :] = W0
    elif init_eig:  # ive.py:110-123: un-conjugated leading eigenvector of general eig
        vals, vecs = np.linalg.eig(Cx)
        lead = np.argmax(vals, axis=1)
        w[:, :, 0] = vecs[np.arange(F), :, lead]
    else:
        w[:, 0] = 1.0  # ive.py:125-127

    def a_from_w(mask):  # ive.py:132-135
        v_new = Cx[mask] @ w[mask]
        lam_w = 1.0 / np.real(herm(w[mask]) @ v_new)
        a[mask] = lam_w * v_new

    def w_from_a(mask):  # ive.py:137-140 (lambda_a refreshed for every bin, w only where masked)
        v_new = Cx_inv @ a
        lambda_a[:] = 1.0 / np.real(herm(a) @ v_new)
        w[mask] = lambda_a[mask] * v_new[mask]

    def switching():  # ive.py:142-161
        a_n = a / a[:, :1, :1]
        b_n = Cx @ a_n
        lmb = b_n[:, :1, :1].copy()
        b_n = b_n / lmb
        p1 = np.linalg.norm(a_n - b_n, axis=(1, 2)) / Cx_norm
        Cbb = lmb * (b_n @ herm(b_n)) / np.linalg.norm(b_n, axis=(1, 2), keepdims=True) ** 2
        p2 = np.linalg.norm(Cx - Cbb, axis=(1, 2))
        kappa = p1 * p2 / np.sqrt(M)
        return kappa >= 0.1, kappa < 0.1  # (do_a, do_w)

    a_from_w(np.ones(F, dtype=bool))  # ive.py:168
    if update == "mix":  # ive.py:170-175
        do_w = np.zeros(F, dtype=bool)
        do_a = np.ones(F, dtype=bool)
    else:
        do_w = np.ones(F, dtype=bool)
        do_a = np.zeros(F, dtype=bool)

    Xf = np.ascontiguousarray(X.swapaxes(0, 1))

    for epoch in range(n_iter):
        if update == "switching" and epoch % 10 == 0:  # ive.py:187-188
            do_a, do_w = switching()

        Y = demix(Xf, w)  # (F, T, 1), ive.py:191

        if callback is not None and epoch % 100 == 0:  # ive.py:194-200
            Y_tmp = Y.swapaxes(0, 1)
            if proj_back:
                z = projection_back(Y_tmp, X[:, :, 0])
                callback(Y_tmp * np.conj(z[None, :, :]))
            else:
                callback(Y_tmp)

        r2 = np.einsum("ftk,ftk->tk", np.conj(Y), Y).real
        if model == "laplace":  # ive.py:204-205
            r = np.sqrt(r2) / np.sqrt(F)
        elif model == "gauss":  # ive.py:207-208
            r = r2 / F
        else:
            r = np.zeros_like(r2)
        r = np.where(r < EPS_R, EPS_R, r)  # ive.py:212-213
        r_inv = 1.0 / r

        V = weighted_covariance(Xf, r_inv[:, 0])  # (F, M, M)
        Vw = V @ w  # (F, M, 1)
        zeta = herm(w) @ Vw  # (F, 1, 1) == (1/T) sum_t |y|^2 / r, ive.py:218
        x_psi = Vw / zeta  # ive.py:220

        # ive.py:222-232: masked w-step / a-step
        delta[do_w] = a[do_w] - x_psi[do_w]
        w[do_w] += step_size * delta[do_w]
        delta[do_a] = w[do_a] - (Cx_inv[do_a] @ x_psi[do_a]) * lambda_a[do_a]
        a[do_a] += step_size * delta[do_a]

        a_from_w(do_w)  # ive.py:235-236
        w_from_a(do_a)

        max_delta = np.max(np.linalg.norm(delta, axis=(1, 2)))  # ive.py:238-241
        if max_delta < tol:
            break

    Y = np.ascontiguousarray(demix(Xf, w).swapaxes(0, 1))  # ive.py:244-246
    if proj_back:
        z = projection_back(Y, X[:, :, 0])
        Y *= np.conj(z[None, :, :])
    if return_filters:
        return Y, w
    return Y
