"""STFT analysis / synthesis on the GPU and the "audio in -> audio out" pipeline around the demixing loop.

Mirrors what the reference's drivers do with pyroomacoustics either side of the algorithm call
(``overiva_oneshot.py:156-158,293-295,371-379``; ``overiva_sim.py:98-100,206-207,213-218``):

    win_a = hann(framesize);  win_s = compute_synthesis_window(win_a, framesize // 2)
    X = analysis(mix.T, framesize, framesize // 2, win=win_a)          # (n_frames, n_freq, n_mics)
    Y = overiva(X, ...)
    y = synthesis(Y, framesize, framesize // 2, win=win_s)              # (n_samples', n_src)

``analysis`` / ``synthesis`` / ``hann`` / ``compute_synthesis_window`` keep those names and argument order.
:func:`separate` runs the three steps without leaving the device: the analysis kernel writes the spectra
directly in the grouped layout the loop streams (no (T,F,M) array, no relayout pass), so only audio crosses
PCIe -- half the bytes of the complex128 STFT in each direction.

Framing: no implicit padding, ``n_frames = (N + pad_front + pad_back - L)//hop + 1``; ``pad_front`` /
``pad_back`` zeros are virtual (never materialised).  All transforms are hand-written fp64 kernels
(``csrc/stft.cu``); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import core

_TW_CACHE = {}


def hann(n):
    """Periodic Hann window, ``0.5 (1 - cos(2 pi i / n))`` (``pra.hann(framesize)``, overiva_oneshot.py:157)."""
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(int(n)) / int(n)))


def compute_synthesis_window(win_a, hop):
    """Matched synthesis window (``pra.transform.compute_synthesis_window``, overiva_oneshot.py:158):
    ``win_a`` divided by the sum of its squared copies shifted by every multiple of ``hop``."""
    win_a = np.asarray(win_a, dtype=np.float64)
    n = win_a.shape[0]
    hop = int(hop)
    norm = np.zeros(n)
    for shift in range(-((n - 1) // hop) * hop, n, hop):
        lo, hi = max(0, shift), min(n, n + shift)
        norm[lo:hi] += win_a[lo - shift : hi - shift] ** 2
    return win_a / norm


def num_frames(n_samples, L_, hop, pad_front=0, pad_back=0):
    return int(L.load().oiva_stft_num_frames(int(n_samples), int(L_), int(hop), int(pad_front), int(pad_back)))


def _twiddles(frame_len, device):
    key = (int(frame_len), torch.device(device).index)
    tw = _TW_CACHE.get(key)
    if tw is None:
        nbytes = L.load().oiva_stft_twiddle_bytes(int(frame_len))
        if nbytes == 0:
            raise ValueError("frame length %d must be a power of two in 8..8192" % frame_len)
        tw = torch.empty(nbytes // 16, dtype=torch.complex128, device=device)
        L.check(L.load().oiva_stft_twiddles(core._ptr(tw), int(frame_len), core._stream_ptr(device)),
                "oiva_stft_twiddles")
        _TW_CACHE[key] = tw
    return tw


def _window(win, frame_len, device):
    if win is None:
        return None
    w = torch.as_tensor(np.asarray(win, dtype=np.float64)) if not isinstance(win, torch.Tensor) else win
    if w.numel() != frame_len:
        raise ValueError("window has %d samples, frame length is %d" % (w.numel(), frame_len))
    return w.to(device=device, dtype=torch.float64).contiguous()


class _Audio:
    """Real audio (N,), (N, M) or (B, N, M) in any strides -> a CUDA tensor viewed as (B, N, M)."""

    def __init__(self, x, device=None):
        self.kind = "numpy"
        if isinstance(x, torch.Tensor):
            self.kind = "cuda" if x.is_cuda else "cpu"
            t = x
        else:
            x = np.asarray(x)
            if np.iscomplexobj(x):
                raise TypeError("audio must be real")
            if x.dtype not in (np.float64, np.float32):
                x = x.astype(np.float64)
            t = torch.from_numpy(x)
        if t.dtype not in (torch.float64, torch.float32):
            raise TypeError("audio must be float32 or float64, got %s" % t.dtype)
        if t.ndim not in (1, 2, 3):
            raise ValueError("audio must have shape (n_samples,), (n_samples, n_chan) or (n_batch, n_samples, n_chan)")
        self.ndim = t.ndim
        self.device = core._require_cuda(t.device if t.is_cuda else device)
        d = t.to(self.device, non_blocking=True)
        if d.ndim == 1:
            d = d[:, None]
        if d.ndim == 2:
            d = d[None]
        self.dev = d
        self.dtype = t.dtype

    def give_back(self, t):
        if self.kind == "cuda":
            return t
        out = t.cpu()
        return out.numpy() if self.kind == "numpy" else out


def _analysis_launch(a, frame_len, hop, win, pad_front, n_frames, out, grouped, cdtype):
    lib = L.load()
    d = a.dev
    B, N, M = d.shape
    sb, sn, sc = d.stride()
    # keep the window / twiddle tensors referenced until the launch has been queued: a temporary freed before
    # the launch could be handed to the next allocation on this stream
    wd, tw = _window(win, frame_len, a.device), _twiddles(frame_len, a.device)
    L.check(lib.oiva_stft_analysis(core._ptr(d), int(d.dtype == torch.float32), sb, sn, sc, N, int(pad_front),
                                   core._ptr(wd), core._ptr(tw),
                                   out if isinstance(out, C.c_void_p) else core._ptr(out), int(grouped), B, n_frames, M,
                                   int(frame_len), int(hop), L.C64 if cdtype == torch.complex64 else L.C128,
                                   core._stream_ptr(a.device)), "oiva_stft_analysis")


def analysis(x, L_, hop, win=None, pad_front=0, pad_back=0, dtype=None):
    """``pra.transform.analysis(x, L, hop, win=win)`` (overiva_oneshot.py:293-295): x (N,), (N, M) or
    (B, N, M) real -> X (T, F), (T, F, M) or (B, T, F, M) complex with ``F = L//2 + 1`` and
    ``X[t] = rfft(win * x[t*hop - pad_front : ... + L])``.  complex128 unless x is float32 (then complex64)
    or ``dtype`` says otherwise.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    a = _Audio(x)
    B, N, M = a.dev.shape
    T = num_frames(N, L_, hop, pad_front, pad_back)
    if T <= 0:
        raise ValueError("signal of %d samples is shorter than one frame of %d" % (N + pad_front + pad_back, L_))
    if M > 16:
        raise ValueError("at most 16 channels are supported, got %d" % M)
    cdtype = dtype or (torch.complex64 if a.dtype == torch.float32 else torch.complex128)
    if not isinstance(cdtype, torch.dtype):
        cdtype = torch.complex64 if np.dtype(cdtype) == np.complex64 else torch.complex128
    with torch.cuda.device(a.device):
        X = torch.empty((B, T, L_ // 2 + 1, M), dtype=cdtype, device=a.device)
        _analysis_launch(a, L_, hop, win, pad_front, T, X, 0, cdtype)
        if a.ndim == 1:
            X = X[0, :, :, 0]
        elif a.ndim == 2:
            X = X[0]
        return a.give_back(X)


def _synthesis_dev(Yd, frame_len, hop, win, out_dtype=torch.float64):
    """Yd (B, T, F, K) contiguous CUDA complex -> (B, (T-1)*hop + L, K) real CUDA tensor."""
    lib = L.load()
    B, T, F, K = Yd.shape
    if F != frame_len // 2 + 1:
        raise ValueError("X has %d bins, frame length %d needs %d" % (F, frame_len, frame_len // 2 + 1))
    dev = Yd.device
    scratch = torch.empty(lib.oiva_stft_scratch_bytes(B, T, K, frame_len), dtype=torch.uint8, device=dev)
    y = torch.empty((B, (T - 1) * hop + frame_len, K), dtype=out_dtype, device=dev)
    wd, tw = _window(win, frame_len, dev), _twiddles(frame_len, dev)  # referenced until the launch is queued
    L.check(lib.oiva_stft_synthesis(core._ptr(Yd), core._ptr(wd), core._ptr(tw), core._ptr(scratch), core._ptr(y),
                                    int(out_dtype == torch.float32), B, T, K, int(frame_len), int(hop),
                                    L.C64 if Yd.dtype == torch.complex64 else L.C128, core._stream_ptr(dev)),
            "oiva_stft_synthesis")
    return y


def synthesis(X, L_, hop, win=None):
    """``pra.transform.synthesis(X, L, hop, win=win)`` (overiva_oneshot.py:371-379): X (T, F), (T, F, K) or
    (B, T, F, K) complex -> y of ``(T-1)*hop + L`` samples, shaped (N,), (N, K) or (B, N, K); float64 for
    complex128 input, float32 for complex64."""
    nd = getattr(X, "ndim", 0)
    if nd not in (2, 3, 4):
        raise ValueError("X must have shape (T, F), (T, F, K) or (B, T, F, K)")
    if nd == 2:
        Xv = X[:, :, None][None]
    elif nd == 3:
        Xv = X[None]
    else:
        Xv = X
    inp = core._Input(Xv)
    with torch.cuda.device(inp.device):
        y = _synthesis_dev(inp.dev, int(L_), int(hop), win,
                           torch.float32 if inp.dtype == torch.complex64 else torch.float64)
        if nd == 2:
            y = y[0, :, 0]
        elif nd == 3:
            y = y[0]
        return inp.give_back(y)


_ALGOS = ("overiva", "auxiva", "auxiva_pca", "ogive", "ilrma")


def separate_batch(mix, n_src=None, n_iter=20, framesize=4096, hop=None, win_a=None, win_s=None, model="laplace",
                   init_eig=False, proj_back=True, pad_front=0, pad_back=0, chunk=32, out=None,
                   dtype=torch.complex128):
    """Host-resident batch of mixtures, audio in -> audio out: mix (B, N, M) real numpy / CPU tensor (pin it for
    full PCIe speed) -> y (B, N', K).  Chunks of ``chunk`` mixtures stream through the GPU with the H2D copy of
    chunk i+1, STFT + loop + iSTFT of chunk i and the D2H copy of chunk i-1 overlapped on three streams -- the
    audio counterpart of ``overiva_batch`` for host inputs.  Only M*N real samples per mixture go up and K*N' come
    back (half the bytes of the complex128 spectra each way).  ``out``: optional preallocated pinned (B, N', K)."""
    hop = framesize // 2 if hop is None else int(hop)
    win_a = hann(framesize) if win_a is None else win_a
    win_s = compute_synthesis_window(win_a, hop) if win_s is None else win_s
    kind = "cpu" if isinstance(mix, torch.Tensor) else "numpy"
    xh = mix if kind == "cpu" else torch.from_numpy(np.ascontiguousarray(mix))
    if xh.is_cuda or xh.ndim != 3 or xh.dtype not in (torch.float64, torch.float32):
        raise ValueError("mix must be a host (n_batch, n_samples, n_mics) float32/float64 array")
    xh = xh.contiguous()
    B, N, M = xh.shape
    K = M if n_src is None else int(n_src)
    if not (1 <= K <= M):
        raise ValueError("n_src=%d must be in 1..n_chan=%d" % (K, M))
    T = num_frames(N, framesize, hop, pad_front, pad_back)
    if T <= 0:
        raise ValueError("signal of %d samples is shorter than one frame of %d" % (N + pad_front + pad_back, framesize))
    F = framesize // 2 + 1
    n_out = (T - 1) * hop + framesize
    out_dtype = torch.float32 if dtype == torch.complex64 else torch.float64
    dev = core._require_cuda()
    code = core._model_code(model)
    with torch.cuda.device(dev):
        yh = out if out is not None else torch.empty((B, n_out, K), dtype=out_dtype, pin_memory=True)
        main = torch.cuda.current_stream(dev)
        owned = core._PIPE_LOCK.acquire(blocking=False)  # a concurrent caller gets private, uncached state
        s_in, s_cmp, s_out = core._pipe_streams(dev) if owned else (torch.cuda.Stream(dev), torch.cuda.Stream(dev),
                                                                     torch.cuda.Stream(dev))
        chunk = min(int(chunk), B)
        lib = L.load()

        def build():  # device slots, scratch and plans of this pipeline shape: kept until another shape comes along
            with torch.cuda.stream(s_cmp):
                return {
                    "xd": [torch.empty((chunk, N, M), dtype=xh.dtype, device=dev) for _ in range(2)],
                    "yd": [torch.empty((chunk, n_out, K), dtype=out_dtype, device=dev) for _ in range(2)],
                    "Yd": torch.empty((chunk, T, F, K), dtype=dtype, device=dev),
                    "scratch": torch.empty(lib.oiva_stft_scratch_bytes(chunk, T, K, framesize), dtype=torch.uint8,
                                           device=dev),
                    "plans": {},
                }

        key = (chunk, N, M, K, T, int(framesize), hop, code, dtype, xh.dtype, torch.device(dev).index)
        try:
            state = core._pipe_state("audio", key, build) if owned else build()
        except BaseException:
            if owned:
                core._PIPE_LOCK.release()
            raise
        xd, yd, Yd, scratch, plans = state["xd"], state["yd"], state["Yd"], state["scratch"], state["plans"]
        status_dev = torch.zeros((B,), dtype=torch.int32, device=dev)  # one status word per mixture
        for st in (s_in, s_cmp, s_out):
            st.wait_stream(main)
        try:
            with torch.cuda.stream(s_cmp):
                wa, ws, tw = _window(win_a, framesize, dev), _window(win_s, framesize, dev), _twiddles(framesize, dev)
            ev_in, ev_cmp, ev_out = [None] * 2, [None] * 2, [None] * 2
            b0 = prev_nb = 0
            for i, nb in enumerate(core._chunk_schedule(B, chunk)):  # (small chunks at both ends: less fill / drain)
                b0 += prev_nb
                prev_nb = nb
                slot = i % 2
                with torch.cuda.stream(s_in):
                    if ev_cmp[slot] is not None:
                        s_in.wait_event(ev_cmp[slot])
                    xd[slot][:nb].copy_(xh[b0 : b0 + nb], non_blocking=True)
                    ev_in[slot] = s_in.record_event()
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[slot])
                    if ev_out[slot] is not None:
                        s_cmp.wait_event(ev_out[slot])
                    if nb not in plans:
                        plans[nb] = core.DemixPlan(nb, T, F, M, K, code, dtype, dev)
                    plan = plans[nb]
                    plan.reset_status()
                    st = core._stream_ptr(dev)
                    L.check(lib.oiva_stft_analysis(core._ptr(xd[slot]), int(xh.dtype == torch.float32), N * M, M, 1, N,
                                                   int(pad_front), core._ptr(wa), core._ptr(tw),
                                                   C.c_void_p(plan.samples_ptr), 1, nb, T, M, int(framesize), hop,
                                                   plan.code, st), "oiva_stft_analysis")
                    plan.adopt_samples()
                    plan.init(L.INIT_EIG if init_eig else L.INIT_EYE)
                    plan.iterate(int(n_iter))
                    plan.output(proj_back, out=Yd[:nb])
                    L.check(lib.oiva_stft_synthesis(core._ptr(Yd), core._ptr(ws), core._ptr(tw), core._ptr(scratch),
                                                    core._ptr(yd[slot]), int(out_dtype == torch.float32), nb, T, K,
                                                    int(framesize), hop, plan.code, st), "oiva_stft_synthesis")
                    status_dev[b0 : b0 + nb].copy_(plan.status_words, non_blocking=True)
                    ev_cmp[slot] = s_cmp.record_event()
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[slot])
                    yh[b0 : b0 + nb].copy_(yd[slot][:nb], non_blocking=True)
                    ev_out[slot] = s_out.record_event()
            for st in (s_in, s_cmp, s_out):
                main.wait_stream(st)
            status = status_dev.cpu().numpy()  # (synchronises the caller's stream)
            main.synchronize()
            core.raise_for_status(status)
        finally:
            if owned:
                core._PIPE_LOCK.release()
    return yh.numpy() if kind == "numpy" else yh


def separate(mix, algo="overiva", n_src=None, n_iter=20, framesize=4096, hop=None, win_a=None, win_s=None,
             model="laplace", init_eig=False, proj_back=True, W0=None, pad_front=0, pad_back=0,
             return_filters=False, dtype=torch.complex128, **algo_kwargs):
    """Audio in -> audio out, entirely on the device: the body of the reference's demo driver
    (``overiva_oneshot.py:293-379``: analysis, algorithm call, synthesis) in one call.

    mix: (n_samples, n_mics) real, or (n_batch, n_samples, n_mics) for independent mixtures of one shape
    (``overiva`` / ``auxiva`` only).  Returns y (n_samples', n_src) [or (n_batch, n_samples', n_src)] with
    ``n_samples' = (n_frames - 1)*hop + framesize`` [and the demixing filters W if ``return_filters``].
    Defaults follow the drivers: Hann analysis window, matched synthesis window, hop = framesize // 2.
    For ``overiva`` / ``auxiva`` the spectra are produced directly in the loop's grouped layout."""
    if algo not in _ALGOS:
        raise ValueError("No such algorithm {}".format(algo))  # overiva_oneshot.py:355
    hop = framesize // 2 if hop is None else int(hop)
    win_a = hann(framesize) if win_a is None else win_a
    win_s = compute_synthesis_window(win_a, hop) if win_s is None else win_s
    a = _Audio(mix)
    if a.ndim == 1:
        raise ValueError("mix must have shape (n_samples, n_mics) or (n_batch, n_samples, n_mics)")
    B, N, M = a.dev.shape
    T = num_frames(N, framesize, hop, pad_front, pad_back)
    if T <= 0:
        raise ValueError("signal of %d samples is shorter than one frame of %d" % (N + pad_front + pad_back, framesize))
    F = framesize // 2 + 1
    cdtype = dtype
    out_dtype = torch.float32 if cdtype == torch.complex64 else torch.float64
    with torch.cuda.device(a.device):
        if algo in ("overiva", "auxiva"):
            if algo_kwargs:
                raise TypeError("unexpected keyword argument %r" % sorted(algo_kwargs)[0])
            K = M if (n_src is None or algo == "auxiva") else int(n_src)
            if not (1 <= K <= M):
                raise ValueError("n_src=%d must be in 1..n_chan=%d" % (K, M))
            if M > 16:
                raise ValueError("at most 16 channels are supported, got %d" % M)
            plan = core._acquire_plan(B, T, F, M, K, core._model_code(model), cdtype, a.device)
            try:
                _analysis_launch(a, framesize, hop, win_a, pad_front, T, C.c_void_p(plan.samples_ptr), 1, cdtype)
                plan.adopt_samples()
                if W0 is not None:
                    plan.init(L.INIT_W0, core._prepare_W0(W0, B, F, M, K, a.device))
                else:
                    plan.init(L.INIT_EIG if init_eig else L.INIT_EYE)
                plan.iterate(int(n_iter))
                Y = plan.output(proj_back)
                W = plan.filters() if return_filters else None
                plan.raise_on_failure()
            finally:
                core._release_plan(plan)
            y = _synthesis_dev(Y, framesize, hop, win_s, out_dtype)
        else:
            if a.ndim == 3:
                raise ValueError("%s takes one mixture at a time" % algo)
            X = torch.empty((1, T, F, M), dtype=cdtype, device=a.device)
            _analysis_launch(a, framesize, hop, win_a, pad_front, T, X, 0, cdtype)
            if algo == "auxiva_pca":
                if return_filters:
                    raise TypeError("auxiva_pca does not support return_filters=True")
                Y = core.auxiva_pca(X[0], n_src, n_iter=n_iter, proj_back=proj_back, W0=W0, model=model,
                                    init_eig=init_eig, **algo_kwargs)
                W = None
            elif algo == "ilrma":  # overiva_oneshot.py:331-339
                from .ilrma import ilrma

                res = ilrma(X[0], n_src, n_iter=n_iter, proj_back=proj_back, W0=W0, return_filters=return_filters,
                            **algo_kwargs)
                Y, W = res if return_filters else (res, None)
                W = W[None] if W is not None else None
            else:
                res = core.ogive(X[0], n_iter=n_iter, proj_back=proj_back, W0=W0, model=model, init_eig=init_eig,
                                 return_filters=return_filters, **algo_kwargs)
                Y, W = res if return_filters else (res, None)
                W = W[None] if W is not None else None
            y = _synthesis_dev(Y[None].contiguous(), framesize, hop, win_s, out_dtype)
        if a.ndim == 2:
            y = y[0]
            W = W[0] if W is not None else None
        y = a.give_back(y)
        if return_filters:
            return y, a.give_back(W)
        return y
