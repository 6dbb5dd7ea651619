"""ILRMA on the GPU: drop-in for ``pyroomacoustics.bss.ilrma`` as the reference drivers call it
(``overiva_oneshot.py:331-339``: ``pra.bss.ilrma(X, n_iter=n_iter, n_components=2, proj_back=True, callback=...)``;
``overiva_sim.py:309-311``).

Determined blind source separation with a low-rank (NMF) spectrogram model per source; the demix, weighted-covariance
and IP-sweep kernels are the ones ``overiva`` runs (``csrc/stream.cuh``, ``cov.cuh`` with one weight per (source, frame,
bin), ``solve_tpb.cuh`` / ``solve.cu``), the NMF updates are ``csrc/ilrma.cu``.  pyroomacoustics is third-party and absent
from the reference tree, so the arithmetic is restated from the published algorithm (see ``oracle/ilrma_oracle.py``:
parity unpinned); conventions follow this package: ``W`` is (n_freq, n_chan, n_src) with the demixing vectors in the
columns, all device arithmetic is fp64 (complex64 input is widened on the device and the result narrowed back).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import core

_EPS = 1e-15  # the clamp of the NMF factors (pyroomacoustics' machine_epsilon-sized floor)


def ilrma(X, n_src=None, n_iter=20, proj_back=False, W0=None, n_components=2, return_filters=0, callback=None,
          T0=None, V0=None):
    """X: (n_frames, n_freq, n_chan) complex (numpy array / CPU tensor / CUDA tensor) -> Y (n_frames, n_freq, n_src) of X's
    type and dtype [, W (n_freq, n_chan, n_src)].  ``n_src`` must equal ``n_chan`` (determined), ``n_chan <= 8``,
    ``n_components <= 8``.  ``T0`` (n_src, n_freq, n_components) / ``V0`` (n_src, n_frames, n_components): initial NMF
    factors; by default drawn as ``0.1 + 0.9 * np.random.rand(...)`` from numpy's global generator, T first, then V.
    ``callback(Y)`` is called every 10th epoch before that epoch's update, with the projected-back estimate when
    ``proj_back``."""
    if getattr(X, "ndim", 0) != 3:
        raise ValueError("X must have shape (n_frames, n_freq, n_chan)")
    inp = core._Input(X)
    T, F, M = inp.dev.shape
    K = M if n_src is None else int(n_src)
    if K != M:
        raise ValueError("ILRMA is a determined algorithm: n_src (%d) must equal the number of channels (%d)" % (K, M))
    if M > 8:
        raise ValueError("ilrma: at most 8 channels are supported, got %d" % M)
    Lc = int(n_components)
    if not (1 <= Lc <= 8):
        raise ValueError("ilrma: n_components must be in 1..8, got %d" % Lc)
    dev = inp.device
    lib = L.load()
    with torch.cuda.device(dev):
        st = core._stream_ptr(dev)
        Xd = inp.dev[None].to(torch.complex128).contiguous()
        T0d = torch.as_tensor(np.asarray(T0, dtype=np.float64) if T0 is not None
                              else 0.1 + 0.9 * np.random.rand(K, F, Lc)).to(dev).contiguous()
        V0d = torch.as_tensor(np.asarray(V0, dtype=np.float64) if V0 is not None
                              else 0.1 + 0.9 * np.random.rand(K, T, Lc)).to(dev).contiguous()
        if tuple(T0d.shape) != (K, F, Lc) or tuple(V0d.shape) != (K, T, Lc):
            raise ValueError("T0 must be (n_src, n_freq, n_components) and V0 (n_src, n_frames, n_components)")
        plan = core._acquire_plan(1, T, F, M, K, L.MODEL_NONE, torch.complex128, dev)
        try:
            plan.reset_status()
            plan.load(Xd)
            if W0 is not None:
                plan.init(L.INIT_W0, core._prepare_W0(W0, 1, F, M, K, dev))
            else:
                plan.init(L.INIT_EYE)
            NG, Tp = lib.oiva_bin_groups(F), lib.oiva_frame_pitch(T)
            G = NG
            arr = lambda which: C.c_void_p(lib.oiva_plan_array(plan.h, which))  # noqa: E731
            Wg, Cg, Vg, r2part, zs = arr(0), arr(1), arr(2), arr(3), arr(5)
            scratch, scratch_bytes = arr(4), lib.oiva_plan_scratch_bytes(plan.h)
            xg = C.c_void_p(plan.samples_ptr)
            f64 = dict(dtype=torch.float64, device=dev)
            Pg = torch.empty(G * K * Tp * 32, **f64)
            iRg = torch.empty(G * K * Tp * 32, **f64)
            Tg = torch.empty(G * K * Lc * 32, **f64)
            Vn = torch.empty(K * Tp * Lc, **f64)
            Vpart = torch.empty(lib.oiva_ilrma_vpart_bytes(1, T, F, K, Lc) // 8, **f64)
            lam = torch.ones(K, **f64)
            P = core._ptr
            L.check(lib.oiva_ilrma_set_model(P(T0d), P(V0d), P(Tg), P(Vn), P(iRg), 1, T, F, K, Lc, st), "oiva_ilrma_set_model")

            def power():
                L.check(lib.oiva_demix_power_full(xg, Wg, M, 1, r2part, P(Pg), 1, T, F, M, K, L.C128, st),
                        "oiva_demix_power_full")

            def current_output():
                """the reference's Y at this point: the last demix (made before the last scale normalisation of W),
                projected back if asked"""
                Y = torch.empty((1, T, F, K), dtype=torch.complex128, device=dev)
                if proj_back:  # (the least-squares scale absorbs the normalisation)
                    L.check(lib.oiva_demix_output_grouped(xg, Wg, Cg, zs, P(Y), 1, T, F, M, K, L.C128, st),
                            "oiva_demix_output_grouped")
                else:
                    L.check(lib.oiva_ilrma_fill_scale(P(lam), zs, 1, F, K, 1, st), "oiva_ilrma_fill_scale")
                    L.check(lib.oiva_demix_output_scaled(xg, Wg, zs, P(Y), 1, T, F, M, K, L.C128, st),
                            "oiva_demix_output_scaled")
                return Y[0].to(inp.dev.dtype)

            def epochs(n):  # nmf -> covariance with per-bin weights -> sweep -> demix + powers -> rescale, n times
                L.check(lib.oiva_ilrma_iterate(xg, Wg, Vg, C.c_void_p(lib.oiva_plan_cov(plan.h)), Cg, r2part, scratch,
                                               scratch_bytes, P(Pg), P(iRg), P(Tg), P(Vn), P(Vpart), P(lam),
                                               C.c_void_p(lib.oiva_plan_status_ptr(plan.h)), 1, T, F, M, Lc, _EPS, int(n), st),
                        "oiva_ilrma_iterate")

            power()
            n_iter = int(n_iter)
            if callback is None:
                epochs(n_iter)
            else:  # every 10th epoch, before its update (the cadence of overiva.py:142)
                epoch = 0
                while epoch < n_iter:
                    if epoch % 10 == 0:
                        plan.raise_on_failure()
                        callback(inp.give_back(current_output()))
                    step = min(10 - epoch % 10, n_iter - epoch)
                    epochs(step)
                    epoch += step
            Y = current_output()
            W = plan.filters()[0] if return_filters else None
            plan.raise_on_failure()
        finally:
            core._release_plan(plan)
        Yo = inp.give_back(Y)
        if return_filters:
            return Yo, inp.give_back(W, inp.dtype)
        return Yo
