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


def grouped(X):
    """X: (B, T, F, M) numpy complex -> device grouped-sample buffer (uint8 tensor) via oiva_relayout."""
    lib = L.load()
    B, T, F, M = X.shape
    code = code_of(X.dtype)
    nbytes = lib.oiva_grouped_bytes(B, T, F, M, code)
    Xd = to_dev(X)
    Xg = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    L.check(lib.oiva_relayout(P(Xd), P(Xg), B, T, F, M, code, stream()), "oiva_relayout")
    return Xg


def grouped_expected(X):
    """numpy restatement of the grouped layout Xg[b*NG+g][t][c][l] (csrc/common.cuh)."""
    B, T, F, M = X.shape
    NG = (F + 31) // 32
    pad = np.zeros((B, T, NG * 32, M), dtype=X.dtype)
    pad[:, :, :F] = X
    # (B, T, NG, 32, M) -> (B, NG, T, M, 32)
    return np.ascontiguousarray(pad.reshape(B, T, NG, 32, M).transpose(0, 2, 1, 4, 3)).reshape(-1)


def weighted_cov(Xg, phi, B, T, F, M, K, code):
    """phi: (B, K, T) numpy or None -> V (B, F, K, M, M) numpy complex128 (grouped result unpacked)."""
    lib = L.load()
    Tp = lib.oiva_frame_pitch(T)
    phid = None
    if phi is not None:
        ph = np.zeros((B, K, Tp))
        ph[:, :, :T] = phi
        phid = to_dev(ph)
    Vg = torch.full((lib.oiva_grouped_cov_bytes(B, F, M, K) // 8,), float("nan"), dtype=torch.float64, device=dev())
    L.check(lib.oiva_weighted_cov(P(Xg), P(phid), P(Vg), B, T, F, M, K, code, stream()), "oiva_weighted_cov")
    V = torch.empty((B, F, K, M, M), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_unpack_cov(P(Vg), P(V), B, F, M, K, stream()), "oiva_unpack_cov")
    torch.cuda.synchronize()
    return V.cpu().numpy()


def pack_cov(V):
    """V (B, F, K, M, M) numpy -> device grouped lower-triangle buffer Vg[gi][k][e][l]."""
    B, F, K, M, _ = V.shape
    NG = (F + 31) // 32
    NE = M * (M + 1) // 2
    out = np.zeros((B, NG, K, NE, 32), dtype=np.complex128)
    pad = np.zeros((B, NG * 32, K, M, M), dtype=np.complex128)
    pad[:, :F] = V
    pad = pad.reshape(B, NG, 32, K, M, M)
    e = 0
    for i in range(M):
        for j in range(i + 1):
            out[:, :, :, e, :] = pad[:, :, :, :, i, j].transpose(0, 1, 3, 2)
            e += 1
    return to_dev(out)


def demix_power(Xg, W, B, T, F, M, K, code):
    """W: (B, F, M, Wc) numpy complex128 -> r2 (B, K, T) numpy (partials summed with oiva_sum_partials)."""
    lib = L.load()
    Tp = lib.oiva_frame_pitch(T)
    nch = lib.oiva_bin_groups(F)
    Wd = to_dev(W.astype(np.complex128))
    part = torch.full((B, nch, K, Tp), np.nan, dtype=torch.float64, device=dev())
    L.check(lib.oiva_demix_power(P(Xg), P(Wd), W.shape[-1], 0, P(part), B, T, F, M, K, code, stream()),
            "oiva_demix_power")
    r2 = torch.empty((B, K, Tp), dtype=torch.float64, device=dev())
    L.check(lib.oiva_sum_partials(P(part), nch, P(r2), B, T, K, stream()), "oiva_sum_partials")
    torch.cuda.synchronize()
    return r2.cpu().numpy()[:, :, :T], part.cpu().numpy()


def source_model(r2, T, F_total, model):
    """r2 (B, K, T) -> (phi (B,K,T), wscale (B,K))"""
    lib = L.load()
    B, K, _ = r2.shape
    Tp = lib.oiva_frame_pitch(T)
    part = np.zeros((B, 1, K, Tp))
    part[:, 0, :, :T] = r2
    partd = to_dev(part)
    phi = torch.empty((B, K, Tp), dtype=torch.float64, device=dev())
    ws = torch.empty((B, K), dtype=torch.float64, device=dev())
    L.check(lib.oiva_source_model(P(partd), 1, P(phi), P(ws), B, T, K, F_total, model, stream()),
            "oiva_source_model")
    torch.cuda.synchronize()
    return phi.cpu().numpy()[:, :, :T], ws.cpu().numpy()


def ip_update(What, V, Cx, wscale, K, grouped_c=True):
    """What (B,F,M,M), V (B,F,K,M,M), Cx (B,F,M,M), wscale (B,K) or None -> (What', status).
    grouped_c=False withholds the grouped covariance, which selects the row-owner (lane group per bin) sweep."""
    lib = L.load()
    B, F, M, _ = What.shape
    Wd, Vgd, Cd = to_dev(What), pack_cov(V), to_dev(Cx)
    Cgd = pack_cov(Cx[:, :, None]) if grouped_c else None
    wsd = to_dev(wscale) if wscale is not None else None
    st = torch.zeros(B, dtype=torch.int32, device=dev())  # one status word per mixture
    NG = lib.oiva_bin_groups(F)
    Wg = torch.empty((B * NG, M * M, 32), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_group_rows(P(Wd), P(Wg), B, F, M * M, stream()), "oiva_group_rows")
    L.check(lib.oiva_ip_update(P(Wg), P(Vgd), P(Cd), P(Cgd), P(wsd), P(st), B, F, M, K, stream()), "oiva_ip_update")
    Wd.fill_(float("nan"))
    L.check(lib.oiva_ungroup_rows(P(Wg), P(Wd), B, F, M * M, stream()), "oiva_ungroup_rows")
    torch.cuda.synchronize()
    per = st.cpu().numpy()
    return Wd.cpu().numpy(), (int(per[0]) if B == 1 else per)


def init_demix(Cx, K, mode, W0=None, evecs=None):
    lib = L.load()
    R, M, _ = Cx.shape
    Cd = to_dev(Cx)
    W0d = to_dev(W0) if W0 is not None else None
    Ed = to_dev(evecs) if evecs is not None else None
    What = torch.empty((R, M, M), dtype=torch.complex128, device=dev())
    st = torch.zeros(4, dtype=torch.int32, device=dev())
    L.check(lib.oiva_init_demix(P(What), P(Cd), P(W0d), P(Ed), mode, P(st), R, R, M, K, stream()), "oiva_init_demix")
    torch.cuda.synchronize()
    return What.cpu().numpy(), int(st[0].item())


def eigh(Cx, lapack_phase):
    lib = L.load()
    R, M, _ = Cx.shape
    Cd = to_dev(Cx)
    ev = torch.empty((R, M), dtype=torch.float64, device=dev())
    vec = torch.empty((R, M, M), dtype=torch.complex128, device=dev())
    st = torch.zeros(4, dtype=torch.int32, device=dev())
    L.check(lib.oiva_eigh(P(Cd), P(ev), P(vec), P(st), R, R, M, int(lapack_phase), stream()), "oiva_eigh")
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


def demix_output(Xg, Weff, B, T, F, M, K, code):
    lib = L.load()
    Wd = to_dev(Weff)
    dt = torch.complex64 if code == L.C64 else torch.complex128
    Y = torch.full((B, T, F, K), float("nan"), dtype=dt, device=dev())
    L.check(lib.oiva_demix_output(P(Xg), P(Wd), P(Y), B, T, F, M, K, code, stream()), "oiva_demix_output")
    torch.cuda.synchronize()
    return Y.cpu().numpy()


def group_rows(rows, B, F, n_elems):
    """rows (B*F, n_elems) numpy complex128 -> device grouped [B*NG][n_elems][32]."""
    lib = L.load()
    NG = lib.oiva_bin_groups(F)
    Rd = to_dev(rows.astype(np.complex128))
    Gd = torch.empty((B * NG, n_elems, 32), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_group_rows(P(Rd), P(Gd), B, F, n_elems, stream()), "oiva_group_rows")
    return Gd


def ungroup_rows(Gd, B, F, n_elems):
    lib = L.load()
    Rd = torch.full((B * F, n_elems), float("nan"), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_ungroup_rows(P(Gd), P(Rd), B, F, n_elems, stream()), "oiva_ungroup_rows")
    torch.cuda.synchronize()
    return Rd.cpu().numpy()


def init_demix_grouped(Cx, K, W0=None):
    """Cx (B, F, M, M), W0 (B, F, M, K) or None -> (What (B, F, M, M) ungrouped, status (B,), rc)."""
    lib = L.load()
    B, F, M, _ = Cx.shape
    NG = lib.oiva_bin_groups(F)
    Cgd = pack_cov(Cx[:, :, None])
    W0d = to_dev(W0.astype(np.complex128)) if W0 is not None else None
    Wg = torch.full((B * NG, M * M, 32), float("nan"), dtype=torch.complex128, device=dev())
    st = torch.zeros(B, dtype=torch.int32, device=dev())
    rc = lib.oiva_init_demix_grouped(P(Wg), P(Cgd), P(W0d), P(st), B, F, M, K, stream())
    if rc != 0:
        return None, None, rc
    torch.cuda.synchronize()
    # padded bins of a ragged last group must hold zeros
    if F % 32:
        pad = Wg.view(B, NG, M * M, 32)[:, -1, :, F % 32:]
        assert bool((pad == 0).all())
    return ungroup_rows(Wg, B, F, M * M).reshape(B, F, M, M), st.cpu().numpy(), 0


def demix_output_grouped(Xg, What, Cx, B, T, F, M, K, code, proj_back):
    """What (B, F, M, M), Cx (B, F, M, M) numpy -> Y (B, T, F, K) through the single-launch grouped output kernel."""
    lib = L.load()
    Wg = group_rows(What.reshape(B * F, M * M), B, F, M * M)
    Cgd = pack_cov(Cx[:, :, None]) if proj_back else None
    dt = torch.complex64 if code == L.C64 else torch.complex128
    Y = torch.full((B, T, F, K), float("nan"), dtype=dt, device=dev())
    zs = torch.empty((B * lib.oiva_bin_groups(F) * K * 32,), dtype=torch.complex128, device=dev())
    L.check(lib.oiva_demix_output_grouped(P(Xg), P(Wg), P(Cgd), P(zs), P(Y), B, T, F, M, K, code, stream()),
            "oiva_demix_output_grouped")
    torch.cuda.synchronize()
    return Y.cpu().numpy()
