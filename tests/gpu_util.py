"""Helpers for the GPU parity tests: thin numpy-in / numpy-out wrappers around the individual C-ABI
kernels (the same entry points the product path sequences through the plan)."""
import ctypes as C

import numpy as np
import torch

from overiva_b200 import _lib as L


def dev():
    return torch.device("cuda", torch.cuda.current_device())


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def code_of(dtype):
    return L.C64 if dtype in (np.complex64, torch.complex64) else L.C128


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def planar(X):
    """X: (B, T, F, M) numpy complex -> device planar buffer (uint8 tensor) via oiva_relayout."""
    lib = L.load()
    B, T, F, M = X.shape
    code = code_of(X.dtype)
    nbytes = lib.oiva_planar_bytes(B, T, F, M, code)
    Xd = to_dev(X)
    Xp = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    L.check(lib.oiva_relayout(P(Xd), P(Xp), B, T, F, M, code, stream()), "oiva_relayout")
    return Xp


def planar_expected(X):
    """numpy restatement of the planar layout (csrc/common.cuh) for one batch of mixtures."""
    lib = L.load()
    B, T, F, M = X.shape
    code = code_of(X.dtype)
    TT = lib.oiva_tile_frames(T, M, code)
    nT = (T + TT - 1) // TT
    es = 4 if code == L.C64 else 8
    real = np.float32 if code == L.C64 else np.float64
    row_bytes = lib.oiva_planar_bytes(1, T, 1, M, code)
    TL = row_bytes // (2 * M * es) - (nT - 1) * TT
    out = np.zeros((B, F, row_bytes // es), dtype=real)
    for ti in range(nT):
        t0 = ti * TT
        pitch = TT if ti + 1 < nT else TL
        nv = min(TT, T - t0)
        blk = np.zeros((B, F, 2 * M, pitch), dtype=real)
        seg = X[:, t0 : t0 + nv]  # (B, nv, F, M)
        blk[:, :, 0::2, :nv] = seg.real.transpose(0, 2, 3, 1)
        blk[:, :, 1::2, :nv] = seg.imag.transpose(0, 2, 3, 1)
        off = ti * 2 * M * TT
        out[:, :, off : off + 2 * M * pitch] = blk.reshape(B, F, -1)
    return out.reshape(-1)


def weighted_cov(Xp, phi, B, T, F, M, K, code):
    """phi: (B, K, T) numpy or None -> V (B, F, K, M, M) numpy complex128."""
    lib = L.load()
    Tp = lib.oiva_frame_pitch(T, M, code)
    phid = None
    if phi is not None:
        ph = np.zeros((B, K, Tp))
        ph[:, :, :T] = phi
        phid = to_dev(ph)
    V = torch.empty((B, F, K, M, M), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_weighted_cov(P(Xp), P(phid), P(V), B, T, F, M, K, code, stream()), "oiva_weighted_cov")
    torch.cuda.synchronize()
    return V.cpu().numpy()


def demix_power(Xp, W, B, T, F, M, K, code, n_chunks=None):
    """W: (B, F, M, Wc) numpy complex128 -> r2 (B, K, T) numpy (partials summed with oiva_sum_partials)."""
    lib = L.load()
    Tp = lib.oiva_frame_pitch(T, M, code)
    nch = n_chunks or lib.oiva_power_chunks(B, F)
    Wd = to_dev(W.astype(np.complex128))
    part = torch.full((B, nch, K, Tp), np.nan, dtype=torch.float64, device=dev())
    L.check(lib.oiva_demix_power(P(Xp), P(Wd), W.shape[-1], P(part), nch, B, T, F, M, K, code, stream()),
            "oiva_demix_power")
    r2 = torch.empty((B, K, Tp), dtype=torch.float64, device=dev())
    L.check(lib.oiva_sum_partials(P(part), nch, P(r2), B, T, M, K, code, stream()), "oiva_sum_partials")
    torch.cuda.synchronize()
    return r2.cpu().numpy()[:, :, :T], part.cpu().numpy()


def source_model(r2, T, M, F_total, model, code):
    """r2 (B, K, T) -> (phi (B,K,T), wscale (B,K))"""
    lib = L.load()
    B, K, _ = r2.shape
    Tp = lib.oiva_frame_pitch(T, M, code)
    part = np.zeros((B, 1, K, Tp))
    part[:, 0, :, :T] = r2
    partd = to_dev(part)
    phi = torch.empty((B, K, Tp), dtype=torch.float64, device=dev())
    ws = torch.empty((B, K), dtype=torch.float64, device=dev())
    L.check(lib.oiva_source_model(P(partd), 1, P(phi), P(ws), B, T, M, K, F_total, model, code, stream()),
            "oiva_source_model")
    torch.cuda.synchronize()
    return phi.cpu().numpy()[:, :, :T], ws.cpu().numpy()


def ip_update(What, V, Cx, wscale, K):
    """What (B,F,M,M), V (B,F,K,M,M), Cx (B,F,M,M), wscale (B,K) or None -> (What', status)"""
    lib = L.load()
    B, F, M, _ = What.shape
    Wd, Vd, Cd = to_dev(What), to_dev(V), to_dev(Cx)
    wsd = to_dev(wscale) if wscale is not None else None
    st = torch.zeros(4, dtype=torch.int32, device=dev())
    L.check(lib.oiva_ip_update(P(Wd), P(Vd), P(Cd), P(wsd), P(st), B, F, M, K, stream()), "oiva_ip_update")
    torch.cuda.synchronize()
    return Wd.cpu().numpy(), int(st[0].item())


def init_demix(Cx, K, mode, W0=None, evecs=None):
    lib = L.load()
    R, M, _ = Cx.shape
    Cd = to_dev(Cx)
    W0d = to_dev(W0) if W0 is not None else None
    Ed = to_dev(evecs) if evecs is not None else None
    What = torch.empty((R, M, M), dtype=torch.complex128, device=dev())
    st = torch.zeros(4, dtype=torch.int32, device=dev())
    L.check(lib.oiva_init_demix(P(What), P(Cd), P(W0d), P(Ed), mode, P(st), R, M, K, stream()), "oiva_init_demix")
    torch.cuda.synchronize()
    return What.cpu().numpy(), int(st[0].item())


def eigh(Cx, lapack_phase):
    lib = L.load()
    R, M, _ = Cx.shape
    Cd = to_dev(Cx)
    ev = torch.empty((R, M), dtype=torch.float64, device=dev())
    vec = torch.empty((R, M, M), dtype=torch.complex128, device=dev())
    st = torch.zeros(4, dtype=torch.int32, device=dev())
    L.check(lib.oiva_eigh(P(Cd), P(ev), P(vec), P(st), R, M, int(lapack_phase), stream()), "oiva_eigh")
    torch.cuda.synchronize()
    return ev.cpu().numpy(), vec.cpu().numpy(), int(st[0].item())


def projback_filters(W, Cx, K, proj_back):
    lib = L.load()
    R, M, wc = W.shape
    Wd, Cd = to_dev(W), to_dev(Cx)
    Weff = torch.empty((R, M, K), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_projback_filters(P(Wd), wc, P(Cd), P(Weff), R, M, K, int(proj_back), stream()),
            "oiva_projback_filters")
    torch.cuda.synchronize()
    return Weff.cpu().numpy()


def demix_output(Xp, Weff, B, T, F, M, K, code):
    lib = L.load()
    Wd = to_dev(Weff)
    dt = torch.complex64 if code == L.C64 else torch.complex128
    Y = torch.full((B, T, F, K), float("nan"), dtype=dt, device=dev())
    L.check(lib.oiva_demix_output(P(Xp), P(Wd), P(Y), B, T, F, M, K, code, stream()), "oiva_demix_output")
    torch.cuda.synchronize()
    return Y.cpu().numpy()
