"""Python host layer: the reference's entry points on top of the CUDA library.

``overiva``, ``auxiva_pca`` and ``ogive`` keep the exact signatures, argument meaning and return shapes of
``overiva.py:28-38``, ``auxiva_pca.py:30`` and ``ive.py:33-45`` of onolab-tmu/overiva, and its error behaviour with
the deviations listed in INTEGRATION.md ("Error behaviour": ``LinAlgError`` for singular bins like the reference,
also raised when a weighted covariance is not positive definite; NaN results with a ``RuntimeWarning`` where the
reference returns NaNs silently; ``ValueError`` for more than 16 channels or an unknown ``model`` string); ``auxiva`` is ``overiva`` with ``n_src`` omitted (``overiva_oneshot.py:301-309``).
Inputs may be numpy arrays (result: numpy), CPU torch tensors (result: CPU torch tensors, pinned when
the input is pinned) or CUDA torch tensors (result: CUDA tensors, nothing leaves the device).

All arithmetic happens in hand-written sm_100a kernels reached through the C ABI
(``include/overiva_b200.h``); PyTorch only provides device memory, streams and (for the multi-GPU
drivers) ``torch.distributed``.  There is no CPU path: without a CUDA device or without the built
library every function raises.
"""
from __future__ import annotations

import ctypes as C
import threading
import warnings

import numpy as np
import torch

from . import _lib as L

_MODELS = {"laplace": L.MODEL_LAPLACE, "gauss": L.MODEL_GAUSS}
_OGIVE_MODELS = {"laplace": L.MODEL_OGIVE_LAPLACE, "gauss": L.MODEL_OGIVE_GAUSS}


def _require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("overiva_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# Pageable host inputs (numpy arrays) are staged through a cached PINNED buffer: a multi-threaded host copy into
# it, then one DMA transfer, instead of the driver's own bounce-buffer copy of pageable memory (config 1, numpy in ->
# numpy out: 4.9 ms -> see DESIGN.md).  Results go to pinned memory from torch's caching host allocator (a freed
# result's block is reused by the next call, so steady-state calls allocate nothing).
_STAGING = {}
_STAGING_SEEN = {}  # (shape, dtype) of pageable inputs seen once: pinned staging starts with the SECOND call of a shape
_STAGING_LOCK = threading.Lock()  # (allocating pinned memory costs more than one pageable copy saves)
_STAGING_MAX_BYTES = 256 << 20


def _staging_buffer(shape, dtype):
    nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    if nbytes == 0 or nbytes > _STAGING_MAX_BYTES:
        return None
    key = (tuple(shape), dtype)
    buf = _STAGING.get(key)
    if buf is None:
        if key not in _STAGING_SEEN:
            if len(_STAGING_SEEN) > 64:
                _STAGING_SEEN.clear()
            _STAGING_SEEN[key] = True
            return None
        while _STAGING and sum(b.numel() * b.element_size() for b in _STAGING.values()) + nbytes > _STAGING_MAX_BYTES:
            _STAGING.pop(next(iter(_STAGING)))
        buf = torch.empty(shape, dtype=dtype, pin_memory=True)
        _STAGING[key] = buf
    return buf


class _Input:
    """Normalises the three accepted input kinds to a contiguous CUDA tensor and remembers how to
    hand results back in the caller's kind."""

    def __init__(self, X, device=None):
        self.kind = "numpy"
        self.pinned = False
        if isinstance(X, torch.Tensor):
            self.kind = "cuda" if X.is_cuda else "cpu"
            self.pinned = (not X.is_cuda) and X.is_pinned()
            t = X
        else:
            X = np.asarray(X)
            if not np.iscomplexobj(X):
                raise TypeError("X must be a complex STFT array, got dtype %s" % X.dtype)
            if X.dtype not in (np.complex128, np.complex64):
                X = X.astype(np.complex128)
            X = np.ascontiguousarray(X)
            if not X.flags.writeable:  # torch refuses read-only buffers; the input is never written to anyway
                X = X.copy()
            t = torch.from_numpy(X)
        if t.dtype not in (torch.complex128, torch.complex64):
            raise TypeError("X must be complex64 or complex128, got %s" % t.dtype)
        self.device = _require_cuda(t.device if t.is_cuda else device)
        self.dtype = t.dtype
        self.staged = False
        if t.is_cuda or self.pinned:
            self.dev = t.to(self.device, non_blocking=True).contiguous()
            return
        # pageable host memory: stage through the cached pinned buffer (one caller at a time; a concurrent caller
        # falls back to the plain pageable copy)
        if _STAGING_LOCK.acquire(blocking=False):
            try:
                buf = _staging_buffer(t.shape, t.dtype)
                if buf is not None:
                    buf.copy_(t)
                    self.dev = buf.to(self.device, non_blocking=True)
                    torch.cuda.current_stream(self.device).synchronize()  # the buffer is reusable from here on
                    self.staged = True
                    return
            finally:
                _STAGING_LOCK.release()
        self.dev = t.to(self.device).contiguous()

    def give_back(self, t, dtype=None):
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if self.kind == "cuda":
            return t
        if self.pinned or self.staged:
            out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)  # torch's caching host allocator
            out.copy_(t, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        else:
            out = t.cpu()
        return out.numpy() if self.kind == "numpy" else out


class DemixPlan:
    """Thin owner of an ``oiva_plan_t`` plus its workspace tensor (one per (shape, K, model, dtype))."""

    def __init__(self, n_batch, n_frames, n_freq, n_chan, n_src, model, dtype, device, n_freq_total=0):
        self.lib = L.load()
        self.device = torch.device(device)
        self.B, self.T, self.F, self.M, self.K = n_batch, n_frames, n_freq, n_chan, n_src
        self.cdtype = dtype
        self.code = L.C64 if dtype == torch.complex64 else L.C128
        self.desc = L.PlanDesc(n_batch, n_frames, n_freq, n_freq_total, n_chan, n_src, model, self.code, 0)
        h = C.c_void_p()
        L.check(self.lib.oiva_plan_create(C.byref(h), C.byref(self.desc)), "oiva_plan_create")
        self.h = h
        nbytes = self.lib.oiva_plan_workspace_bytes(h)
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        L.check(self.lib.oiva_plan_bind(h, _ptr(self.ws), nbytes), "oiva_plan_bind")
        self.in_use = False
        self._status_off = self.lib.oiva_plan_status_ptr(h) - self.ws.data_ptr()
        self.reset_status()  # the status words accumulate from here on (see raise_on_failure)
        self.Tp = self.lib.oiva_frame_pitch(n_frames)

    def reset_status(self):
        L.check(self.lib.oiva_plan_reset_status(self.h, _stream_ptr(self.device)), "oiva_plan_reset_status")

    @property
    def status_words(self):
        """(B,) int32 view of the per-mixture status words inside the workspace (device tensor)."""
        return self.ws[self._status_off : self._status_off + 4 * self.B].view(torch.int32)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h is not None and getattr(self, "lib", None) is not None:
            self.lib.oiva_plan_destroy(h)

    # -- views into the workspace -----------------------------------------------------------------
    def _view(self, ptr, nbytes, dtype, shape):
        off = ptr - self.ws.data_ptr()
        return self.ws[off : off + nbytes].view(dtype).view(shape)

    @property
    def r2(self):
        """(B, K, Tp) float64: the source-model statistic summed over this plan's bins (all-reduced by the
        frequency-sharded driver between ``power()`` and ``update()``)."""
        n = self.lib.oiva_plan_r2_elems(self.h)
        return self._view(self.lib.oiva_plan_r2(self.h), n * 8, torch.float64, (self.B, self.K, self.Tp))

    @property
    def what(self):
        n = self.B * self.F * self.M * self.M
        return self._view(self.lib.oiva_plan_what(self.h), n * 16, torch.complex128, (self.B, self.F, self.M, self.M))

    @property
    def cov(self):
        n = self.B * self.F * self.M * self.M
        return self._view(self.lib.oiva_plan_cov(self.h), n * 16, torch.complex128, (self.B, self.F, self.M, self.M))

    @property
    def samples_ptr(self):
        return self.lib.oiva_plan_samples(self.h)

    # -- steps ------------------------------------------------------------------------------------
    def load(self, X):
        assert X.is_cuda and X.is_contiguous() and X.dtype == self.cdtype
        assert tuple(X.shape) == (self.B, self.T, self.F, self.M), (tuple(X.shape), (self.B, self.T, self.F, self.M))
        L.check(self.lib.oiva_plan_load(self.h, _ptr(X), _stream_ptr(self.device)), "oiva_plan_load")

    def adopt_samples(self):
        L.check(self.lib.oiva_plan_adopt_samples(self.h, _stream_ptr(self.device)), "oiva_plan_adopt_samples")

    def init(self, mode, W0=None):
        if W0 is not None:
            assert W0.is_cuda and W0.is_contiguous() and W0.dtype == torch.complex128
            assert tuple(W0.shape) == (self.B, self.F, self.M, self.K)
        L.check(self.lib.oiva_plan_init(self.h, mode, _ptr(W0), _stream_ptr(self.device)), "oiva_plan_init")

    def iterate(self, n_iter):
        if n_iter > 0:
            L.check(self.lib.oiva_plan_iterate(self.h, int(n_iter), _stream_ptr(self.device)), "oiva_plan_iterate")

    def power(self):
        L.check(self.lib.oiva_plan_power(self.h, _stream_ptr(self.device)), "oiva_plan_power")

    def update(self):
        L.check(self.lib.oiva_plan_update(self.h, _stream_ptr(self.device)), "oiva_plan_update")

    def output(self, proj_back, out=None):
        if out is None:
            out = torch.empty((self.B, self.T, self.F, self.K), dtype=self.cdtype, device=self.device)
        L.check(self.lib.oiva_plan_output(self.h, int(bool(proj_back)), _ptr(out), _stream_ptr(self.device)),
                "oiva_plan_output")
        return out

    def filters(self):
        W = torch.empty((self.B, self.F, self.M, self.K), dtype=torch.complex128, device=self.device)
        L.check(self.lib.oiva_plan_filters(self.h, _ptr(W), _stream_ptr(self.device)), "oiva_plan_filters")
        return W

    def run(self, X, init_mode, W0, n_iter, proj_back, return_filters=False, out=None):
        """load + init + iterate + output [+ filters] in ONE library call.  -> (Y, W or None)"""
        assert X.is_cuda and X.is_contiguous() and X.dtype == self.cdtype
        assert tuple(X.shape) == (self.B, self.T, self.F, self.M), (tuple(X.shape), (self.B, self.T, self.F, self.M))
        if W0 is not None:
            assert W0.is_cuda and W0.is_contiguous() and W0.dtype == torch.complex128
            assert tuple(W0.shape) == (self.B, self.F, self.M, self.K)
        Y = out if out is not None else torch.empty((self.B, self.T, self.F, self.K), dtype=self.cdtype, device=self.device)
        W = (torch.empty((self.B, self.F, self.M, self.K), dtype=torch.complex128, device=self.device)
             if return_filters else None)
        L.check(self.lib.oiva_plan_run(self.h, _ptr(X), int(init_mode), _ptr(W0), int(n_iter), int(bool(proj_back)),
                                       _ptr(Y), _ptr(W), _stream_ptr(self.device)), "oiva_plan_run")
        return Y, W

    def status(self):
        """OR of the mixtures' status words (synchronises the stream); negative = library error."""
        return self.lib.oiva_plan_status(self.h, _stream_ptr(self.device))

    def status_vector(self):
        """(B,) numpy int32: one status word per mixture (synchronises the stream)."""
        out = (C.c_int * self.B)()
        st = self.lib.oiva_plan_status_vector(self.h, out, _stream_ptr(self.device))
        if st < 0:
            L.check(st, "oiva_plan_status_vector")
        return np.frombuffer(out, dtype=np.int32).copy()

    def raise_on_failure(self):
        """See :func:`raise_for_status`."""
        st = self.status()
        if st < 0:
            L.check(st, "oiva_plan_status")
        if st == 0:
            return
        raise_for_status(self.status_vector() if self.B > 1 else np.array([st], dtype=np.int32))

    def enable_timing(self, on=True):
        L.check(self.lib.oiva_plan_enable_timing(self.h, int(on)), "oiva_plan_enable_timing")

    def read_timing(self):
        """{'cov': (ms, launches), 'power': ..., 'solve': ...} since the last read; synchronise first."""
        ms = (C.c_double * 3)()
        n = (C.c_longlong * 3)()
        L.check(self.lib.oiva_plan_read_timing(self.h, ms, n), "oiva_plan_read_timing")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("cov", "power", "solve"))}

    @property
    def launches(self):
        return int(self.lib.oiva_plan_launch_count(self.h))


def raise_for_status(status):
    """Turn per-mixture status words into the reference's error behaviour.

    * ``STATUS_SINGULAR`` (an exactly zero / NaN pivot in an LU factorisation, or a covariance that is not positive
      definite in the Cholesky-based update): ``numpy.linalg.LinAlgError("Singular matrix")`` -- what
      ``np.linalg.solve`` raises out of ``overiva.py:98,182`` on rank-deficient input.
    * ``STATUS_NONFINITE`` alone (NaN / Inf in the demixing matrices without a singular pivot, e.g. an all-zero
      input where gamma = 0): the reference silently returns NaN arrays (numpy emits RuntimeWarnings); so does this
      implementation, with one ``RuntimeWarning``.
    * ``STATUS_STALLED`` (a hand-over inside the single-launch loop timed out -- a bug or a broken device, never the
      data): ``RuntimeError``.
    Any other value is a corrupted status word and raises ``RuntimeError``."""
    status = np.atleast_1d(np.asarray(status))
    if np.any((status < 0) | (status > 7)):
        raise RuntimeError("corrupt status words %r" % (status[(status < 0) | (status > 7)][:8],))
    if np.any(status & L.STATUS_STALLED):
        raise RuntimeError("the single-launch loop stalled waiting for another thread block (OIVA_STATUS_STALLED); "
                           "the results are invalid -- rerun with OIVA_NO_RESIDENT=1 and report this")
    sing = np.nonzero(status & L.STATUS_SINGULAR)[0]
    if sing.size:
        if status.size == 1:
            raise np.linalg.LinAlgError("Singular matrix")
        raise np.linalg.LinAlgError("Singular matrix (mixture%s %s of %d)" % (
            "s" if sing.size > 1 else "", ", ".join(str(i) for i in sing[:16]) + (" ..." if sing.size > 16 else ""),
            status.size))
    if np.any(status & L.STATUS_NONFINITE):
        warnings.warn("non-finite values in the demixing matrices (the reference returns NaN arrays here as well)",
                      RuntimeWarning, stacklevel=3)


# ---- plan cache -------------------------------------------------------------------------------------
# A plan (workspace + captured CUDA graph of the epoch loop) is reused by later calls of the same shape: for a
# single short mixture the allocation and the graph capture would otherwise cost more than the separation.
# Only small plans are kept (default budget 1 GiB, OVERIVA_B200_PLAN_CACHE_MB); a plan in use is never shared.
_PLAN_CACHE = {}
_PLAN_CACHE_ORDER = []
_PLAN_CACHE_LOCK = threading.Lock()  # the cache bookkeeping is shared by every thread that calls the entry points


def _plan_cache_budget():
    import os

    return int(float(os.environ.get("OVERIVA_B200_PLAN_CACHE_MB", "1024")) * (1 << 20))


def _acquire_plan(B, T, F, M, K, model_code, dtype, device, n_freq_total=0):
    key = (B, T, F, M, K, model_code, dtype, torch.device(device).index, n_freq_total)
    with _PLAN_CACHE_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is not None and not plan.in_use:
            plan.in_use = True
            _PLAN_CACHE_ORDER.remove(key)
            _PLAN_CACHE_ORDER.append(key)
        else:
            plan = None
    if plan is not None:
        plan.reset_status()
        return plan
    plan = DemixPlan(B, T, F, M, K, model_code, dtype, device, n_freq_total)  # (a concurrent caller gets its own)
    plan.in_use = True
    budget = _plan_cache_budget()
    with _PLAN_CACHE_LOCK:
        if key not in _PLAN_CACHE and plan.ws.numel() <= budget // 2:
            _PLAN_CACHE[key] = plan
            _PLAN_CACHE_ORDER.append(key)
            while sum(_PLAN_CACHE[k].ws.numel() for k in _PLAN_CACHE_ORDER) > budget and len(_PLAN_CACHE_ORDER) > 1:
                old = _PLAN_CACHE_ORDER.pop(0)
                del _PLAN_CACHE[old]
    return plan


def _release_plan(plan):
    with _PLAN_CACHE_LOCK:
        plan.in_use = False


def clear_plan_cache():
    """Drop every cached plan (frees their device workspaces)."""
    with _PLAN_CACHE_LOCK:
        _PLAN_CACHE.clear()
        del _PLAN_CACHE_ORDER[:]
    _PIPE_STATE.clear()
    _STAGING.clear()
    _STAGING_SEEN.clear()


# Host pipelines (overiva_batch / stft.separate_batch on host inputs) keep their three streams for the life of the
# process (the caching allocator keys free blocks by stream: fresh streams per call would mean fresh cudaMallocs per
# call) and the device slots + plans of the MOST RECENT pipeline shape, so that repeated calls of one shape -- a sweep,
# a benchmark -- allocate nothing.  ``clear_plan_cache()`` releases them.
_PIPE_STREAMS = {}
_PIPE_STATE = {}
_PIPE_LOCK = threading.Lock()


def _pipe_streams(dev):
    key = torch.device(dev).index
    st = _PIPE_STREAMS.get(key)
    if st is None:
        st = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        _PIPE_STREAMS[key] = st
    return st


def _pipe_state(name, key, build):
    """The cached state of pipeline ``name`` if it was built for ``key``, else ``build()`` (replacing the old one)."""
    cur = _PIPE_STATE.get(name)
    if cur is not None and cur[0] == key:
        return cur[1]
    _PIPE_STATE.pop(name, None)
    state = build()
    _PIPE_STATE[name] = (key, state)
    return state


def _model_code(model, table=_MODELS):
    try:
        return table[model]
    except (KeyError, TypeError):
        # the reference silently leaves r at zero for an unknown model string and returns NaNs;
        # failing loudly is the one deliberate deviation from it
        raise ValueError("unknown source model %r (expected 'laplace' or 'gauss')" % (model,))


def _prepare_W0(W0, B, F, M, K, device):
    """overiva.py:116-117: ``W[:, :, :] = W0`` broadcasts W0 into (F, M, K)."""
    if isinstance(W0, torch.Tensor):
        W0 = W0.detach().cpu().numpy()
    W0 = np.asarray(W0)
    W0 = np.broadcast_to(W0, (B, F, M, K)) if W0.ndim == 4 else np.broadcast_to(W0, (F, M, K))[None].repeat(B, 0)
    return torch.from_numpy(np.ascontiguousarray(W0.astype(np.complex128))).to(device)


def _run_overiva(Xd, n_src, n_iter, proj_back, W0, model, init_eig, return_filters, callback, cb_wrap,
                 n_freq_total=0, status_out=None):
    """Xd: (B, T, F, M) CUDA tensor.  Returns (Y (B,T,F,K), W (B,F,M,K) or None, plan).  With ``status_out`` (a list)
    numerical failures do not raise: the per-mixture status words are appended to it instead."""
    B, T, F, M = Xd.shape
    if n_src is None:
        n_src = M  # overiva.py:83-84
    n_src = int(n_src)
    if not (1 <= n_src <= M):
        raise ValueError("n_src=%d must be in 1..n_chan=%d" % (n_src, M))
    if M > 16:
        raise ValueError("at most 16 channels are supported, got %d" % M)
    plan = _acquire_plan(B, T, F, M, n_src, _model_code(model), Xd.dtype, Xd.device, n_freq_total)
    try:
        return _run_overiva_on(plan, Xd, n_src, n_iter, proj_back, W0, init_eig, return_filters, callback, cb_wrap,
                               status_out)
    finally:
        _release_plan(plan)


def _run_overiva_on(plan, Xd, n_src, n_iter, proj_back, W0, init_eig, return_filters, callback, cb_wrap,
                    status_out=None):
    B, T, F, M = Xd.shape
    if callback is None:  # the whole call in one library round trip
        W0d = _prepare_W0(W0, B, F, M, n_src, Xd.device) if W0 is not None else None
        mode = L.INIT_W0 if W0 is not None else (L.INIT_EIG if init_eig else L.INIT_EYE)
        Y, W = plan.run(Xd, mode, W0d, n_iter, proj_back, return_filters)
        if status_out is not None:
            status_out.append(plan.status_vector())
        else:
            plan.raise_on_failure()
        return Y, W, plan
    plan.load(Xd)
    if W0 is not None:
        plan.init(L.INIT_W0, _prepare_W0(W0, B, F, M, n_src, Xd.device))
    else:
        plan.init(L.INIT_EIG if init_eig else L.INIT_EYE)
    if callback is None:
        plan.iterate(n_iter)
    else:
        # overiva.py:142-148: every 10th epoch, before that epoch's update, with the projected-back estimate
        epoch = 0
        while epoch < n_iter:
            if epoch % 10 == 0:
                plan.raise_on_failure()
                callback(cb_wrap(plan.output(proj_back)))
            step = min(10 - epoch % 10, n_iter - epoch)
            plan.iterate(step)
            epoch += step
    Y = plan.output(proj_back)
    W = plan.filters() if return_filters else None
    if status_out is not None:
        status_out.append(plan.status_vector())
    else:
        plan.raise_on_failure()
    return Y, W, plan


def overiva(X, n_src=None, n_iter=20, proj_back=True, W0=None, model="laplace", init_eig=False,
            return_filters=False, callback=None):
    """Drop-in for ``overiva.overiva`` (overiva.py:28-204).

    X: (n_frames, n_freq, n_chan) complex.  Returns Y (n_frames, n_freq, n_src) of X's dtype and, if
    ``return_filters``, also W (n_freq, n_chan, n_src).
    """
    if getattr(X, "ndim", 0) != 3:
        raise ValueError("X must have shape (n_frames, n_freq, n_chan)")
    inp = _Input(X)
    with torch.cuda.device(inp.device):
        Y, W, _ = _run_overiva(inp.dev[None], n_src, n_iter, proj_back, W0, model, init_eig, return_filters,
                               callback, lambda y: inp.give_back(y[0]))
        Yo = inp.give_back(Y[0])
        if return_filters:
            return Yo, inp.give_back(W[0], inp.dtype)
        return Yo


def auxiva(X, n_iter=20, proj_back=True, W0=None, model="laplace", init_eig=False, return_filters=False,
           callback=None):
    """Determined AuxIVA: ``overiva`` with ``n_src`` omitted (README.md:178-180, overiva_oneshot.py:301-309)."""
    return overiva(X, None, n_iter, proj_back, W0, model, init_eig, return_filters, callback)


def _chunk_schedule(B, chunk):
    """Chunk sizes of a host pipeline over B mixtures.  The first H2D copy and the last loop + D2H copy are the only
    parts of the call that nothing overlaps, so long batches start and end with small chunks (chunk/8, chunk/8, chunk/4,
    chunk/2 and back) and run full chunks in between: 512 mixtures, chunk 32: ~20 ms less fill / drain per call."""
    chunk = max(1, int(chunk))
    if chunk < 8 or B < 6 * chunk:
        return [chunk] * (B // chunk) + ([B % chunk] if B % chunk else [])
    up = [chunk // 8, chunk // 8, chunk // 4, chunk // 2]
    mid = B - 2 * sum(up)
    return up + [chunk] * (mid // chunk) + ([mid % chunk] if mid % chunk else []) + up[::-1]


def _host_pipeline(Xh, n_src, n_iter, proj_back, W0, model, init_eig, return_filters, chunk, out, device,
                   return_status=False):
    """Host-resident batch through the GPU in chunks: H2D of chunk i+1, the loop on chunk i and D2H of chunk
    i-1 run concurrently on three streams (PCIe is full duplex), with two device slots per direction.  The
    whole call then costs about max(PCIe time, compute time) instead of their sum."""
    B, T, F, M = Xh.shape
    K = M if n_src is None else int(n_src)
    if not (1 <= K <= M):
        raise ValueError("n_src=%d must be in 1..n_chan=%d" % (K, M))
    dev = device
    cdt = Xh.dtype
    code = _model_code(model)
    Yh = out if out is not None else torch.empty((B, T, F, K), dtype=cdt, pin_memory=True)
    Wh = torch.empty((B, F, M, K), dtype=torch.complex128, pin_memory=True) if return_filters else None
    W0d = _prepare_W0(W0, B, F, M, K, dev) if W0 is not None else None
    main = torch.cuda.current_stream(dev)
    owned = _PIPE_LOCK.acquire(blocking=False)  # a concurrent caller (another thread) gets private, uncached state
    try:
        return _host_pipeline_locked(Xh, K, n_iter, proj_back, W0d, code, init_eig, return_filters, chunk, Yh, Wh, dev,
                                     main, owned, return_status)
    finally:
        if owned:
            _PIPE_LOCK.release()


def _host_pipeline_locked(Xh, K, n_iter, proj_back, W0d, code, init_eig, return_filters, chunk, Yh, Wh, dev, main, owned,
                          return_status=False):
    B, T, F, M = Xh.shape
    cdt = Xh.dtype
    s_in, s_cmp, s_out = _pipe_streams(dev) if owned else (torch.cuda.Stream(dev), torch.cuda.Stream(dev),
                                                            torch.cuda.Stream(dev))
    n_slots = 2

    def build():
        return {
            "Xd": [torch.empty((chunk, T, F, M), dtype=cdt, device=dev) for _ in range(n_slots)],
            "Yd": [torch.empty((chunk, T, F, K), dtype=cdt, device=dev) for _ in range(n_slots)],
            "Wd": [torch.empty((chunk, F, M, K), dtype=torch.complex128, device=dev) if return_filters else None
                   for _ in range(n_slots)],
            "plans": {},
        }

    key = (chunk, T, F, M, K, code, cdt, torch.device(dev).index, bool(return_filters))
    state = _pipe_state("spectra", key, build) if owned else build()
    Xd, Yd, Wd, plans = state["Xd"], state["Yd"], state["Wd"], state["plans"]
    status_dev = torch.zeros((B,), dtype=torch.int32, device=dev)  # one status word per mixture of the whole batch
    for st in (s_in, s_cmp, s_out):
        st.wait_stream(main)

    def plan_for(nb, slot):
        key = (nb, slot)
        if key not in plans:
            with torch.cuda.stream(s_cmp):
                plans[key] = DemixPlan(nb, T, F, M, K, code, cdt, dev)
        return plans[key]

    ev_in = [None] * n_slots   # H2D of the chunk in slot s finished
    ev_cmp = [None] * n_slots  # compute on slot s finished (X slot reusable, Y slot filled)
    ev_out = [None] * n_slots  # D2H of slot s finished (Y slot reusable)
    b0 = prev_nb = 0
    for i, nb in enumerate(_chunk_schedule(B, chunk)):
        b0 += prev_nb
        prev_nb = nb
        slot = i % n_slots
        with torch.cuda.stream(s_in):
            if ev_cmp[slot] is not None:
                s_in.wait_event(ev_cmp[slot])
            Xd[slot][:nb].copy_(Xh[b0 : b0 + nb], non_blocking=True)
            ev_in[slot] = s_in.record_event()
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[slot])
            if ev_out[slot] is not None:
                s_cmp.wait_event(ev_out[slot])
            plan = plan_for(nb, slot)
            plan.reset_status()
            mode = L.INIT_W0 if W0d is not None else (L.INIT_EIG if init_eig else L.INIT_EYE)
            plan.run(Xd[slot][:nb], mode, W0d[b0 : b0 + nb].contiguous() if W0d is not None else None, n_iter, proj_back,
                     out=Yd[slot][:nb])  # load + init + iterate + output in one library call
            if return_filters:
                L.check(plan.lib.oiva_plan_filters(plan.h, _ptr(Wd[slot]), _stream_ptr(dev)), "oiva_plan_filters")
            status_dev[b0 : b0 + nb].copy_(plan.status_words, non_blocking=True)
            ev_cmp[slot] = s_cmp.record_event()
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[slot])
            Yh[b0 : b0 + nb].copy_(Yd[slot][:nb], non_blocking=True)
            if return_filters:
                Wh[b0 : b0 + nb].copy_(Wd[slot][:nb], non_blocking=True)
            ev_out[slot] = s_out.record_event()
    for st in (s_in, s_cmp, s_out):
        main.wait_stream(st)
    status = status_dev.cpu().numpy()  # (synchronises the caller's stream, which has waited for the three others)
    main.synchronize()
    if return_status:
        return Yh, Wh, status
    raise_for_status(status)
    return Yh, Wh, None


def overiva_batch(X, n_src=None, n_iter=20, proj_back=True, W0=None, model="laplace", init_eig=False,
                  return_filters=False, chunk=32, out=None, return_status=False):
    """Many independent mixtures at once: X (B, n_frames, n_freq, n_chan) -> Y (B, n_frames, n_freq, n_src)
    [, W (B, n_freq, n_chan, n_src)].  The role of the reference's task farm (``overiva_sim.py`` +
    ``rrtools``) for mixtures of one shape; every mixture is processed exactly as ``overiva`` would.

    Host inputs (numpy / CPU tensors) with more than ``chunk`` mixtures are streamed through the GPU in chunks
    with copies and compute overlapped (pin the input for full PCIe speed; ``out`` may be a preallocated pinned
    (B, T, F, K) tensor to receive Y).  Device inputs are processed in one piece.

    Failures are tracked PER MIXTURE.  By default a singular mixture raises ``LinAlgError`` naming the offending
    mixtures (what a plain loop over ``overiva`` would do at the first of them).  With ``return_status=True`` nothing
    is raised: a last return value ``status`` (B,) int32 holds each mixture's status word (0 = fine, bit 0 = singular,
    bit 1 = non-finite) and the caller decides -- the reference's sweep records NaN for the failing mixture only and
    carries on (``overiva_sim.py:334-350``)."""
    if getattr(X, "ndim", 0) != 4:
        raise ValueError("X must have shape (n_batch, n_frames, n_freq, n_chan)")
    is_host = not (isinstance(X, torch.Tensor) and X.is_cuda)
    if is_host and X.shape[0] > chunk:
        kind = "cpu" if isinstance(X, torch.Tensor) else "numpy"
        Xh = X if kind == "cpu" else torch.from_numpy(np.ascontiguousarray(X))
        if Xh.dtype not in (torch.complex128, torch.complex64):
            raise TypeError("X must be complex64 or complex128, got %s" % Xh.dtype)
        dev = _require_cuda()
        with torch.cuda.device(dev):
            Yh, Wh, status = _host_pipeline(Xh.contiguous(), n_src, n_iter, proj_back, W0, model, init_eig,
                                            return_filters, int(chunk), out, dev, return_status)
        if kind == "numpy":
            Yh = Yh.numpy()
            Wh = Wh.numpy().astype(X.dtype) if Wh is not None else None
        elif Wh is not None:
            Wh = Wh.to(Xh.dtype)
        res = (Yh, Wh) if return_filters else (Yh,)
    else:
        inp = _Input(X)
        st = [] if return_status else None
        with torch.cuda.device(inp.device):
            Y, W, _ = _run_overiva(inp.dev, n_src, n_iter, proj_back, W0, model, init_eig, return_filters, None, None,
                                   status_out=st)
            Yo = inp.give_back(Y)
            if out is not None:
                out.copy_(Yo if isinstance(Yo, torch.Tensor) else torch.from_numpy(Yo))
                Yo = out
            res = (Yo, inp.give_back(W, inp.dtype)) if return_filters else (Yo,)
        status = st[0] if return_status else None
    if return_status:
        res = res + (status,)
    return res if len(res) > 1 else res[0]


def auxiva_pca(X, n_src=None, **kwargs):
    """Drop-in for ``auxiva_pca.auxiva_pca`` (auxiva_pca.py:30-92): PCA to ``n_src`` channels, determined
    AuxIVA on the reduced signal, projection back onto the ORIGINAL microphone 0.

    As in the reference, ``proj_back`` must be present in kwargs (it is popped unconditionally,
    auxiva_pca.py:86, and the projection back is always applied), the remaining kwargs go to ``overiva``,
    and ``return_filters=True`` is not supported (the reference fails on it with a TypeError).
    """
    if getattr(X, "ndim", 0) != 3:
        raise ValueError("X must have shape (n_frames, n_freq, n_chan)")
    kwargs = dict(kwargs)
    kwargs.pop("proj_back")  # KeyError if absent, exactly like auxiva_pca.py:86
    if kwargs.pop("return_filters", False):
        raise TypeError("auxiva_pca does not support return_filters=True")
    n_iter = kwargs.pop("n_iter", 20)
    W0 = kwargs.pop("W0", None)
    model = kwargs.pop("model", "laplace")
    init_eig = kwargs.pop("init_eig", False)
    callback = kwargs.pop("callback", None)
    if "n_src" in kwargs or kwargs:
        raise TypeError("overiva() got an unexpected keyword argument %r" % sorted(kwargs)[0])
    inp = _Input(X)
    T, F, M = inp.dev.shape
    K = M if n_src is None else int(n_src)
    if not (1 <= K <= M):
        raise ValueError("n_src=%d must be in 1..n_chan=%d" % (K, M))
    lib = L.load()
    dev = inp.device
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        code = L.C64 if inp.dtype == torch.complex64 else L.C128
        full = DemixPlan(1, T, F, M, K, _model_code(model), inp.dtype, dev)
        full.load(inp.dev[None])
        if K < M:
            # eigh of the input covariance, keep the K principal eigenvectors      auxiva_pca.py:71-81
            evals = torch.empty((F, M), dtype=torch.float64, device=dev)
            evecs = torch.empty((F, M, M), dtype=torch.complex128, device=dev)
            L.check(lib.oiva_eigh(_ptr(full.cov), _ptr(evals), _ptr(evecs), lib.oiva_plan_status_ptr(full.h), F, F, M, 0,
                                  st), "oiva_eigh")
            E = evecs[:, :, M - K :].contiguous()
            red = DemixPlan(1, T, F, K, K, _model_code(model), inp.dtype, dev)
            L.check(lib.oiva_project_rows(full.samples_ptr, _ptr(E), red.samples_ptr, 1, T, F, M, K, code, st),
                    "oiva_project_rows")
            red.adopt_samples()
        else:
            E, red = None, full
        if W0 is not None:
            red.init(L.INIT_W0, _prepare_W0(W0, 1, F, K, K, dev))
        else:
            red.init(L.INIT_EIG if init_eig else L.INIT_EYE)
        epoch = 0
        if callback is None:
            red.iterate(n_iter)
        else:
            while epoch < n_iter:  # inner overiva runs with proj_back=False (auxiva_pca.py:87)
                if epoch % 10 == 0:
                    red.raise_on_failure()
                    callback(inp.give_back(red.output(False)[0]))
                step = min(10 - epoch % 10, n_iter - epoch)
                red.iterate(step)
                epoch += step
        Wr = red.filters()  # (1, F, K, K)
        if E is not None:
            Wfull = torch.empty((F, M, K), dtype=torch.complex128, device=dev)
            L.check(lib.oiva_compose_filters(_ptr(E), _ptr(Wr), _ptr(Wfull), F, M, K, K, st), "oiva_compose_filters")
        else:
            Wfull = Wr[0].contiguous()
        # projection back on the original mic 0 through the full covariance        auxiva_pca.py:89-90
        Weff = torch.empty((F, M, K), dtype=torch.complex128, device=dev)
        L.check(lib.oiva_projback_filters(_ptr(Wfull), K, _ptr(full.cov), _ptr(Weff), F, M, K, 1, st),
                "oiva_projback_filters")
        Y = torch.empty((1, T, F, K), dtype=inp.dtype, device=dev)
        L.check(lib.oiva_demix_output(full.samples_ptr, _ptr(Weff), _ptr(Y), 1, T, F, M, K, code, st),
                "oiva_demix_output")
        red.raise_on_failure()
        if red is not full:
            full.raise_on_failure()
        return inp.give_back(Y[0])


def ogive(X, n_iter=4000, step_size=0.1, tol=1e-3, update="demix", proj_back=True, W0=None, model="laplace",
          init_eig=False, return_filters=False, callback=None):
    """Drop-in for ``ive.ogive`` (ive.py:33-256): orthogonally constrained gradient extraction of ONE source.

    Returns Y (n_frames, n_freq, 1) and, if ``return_filters``, w (n_freq, n_chan, 1).  The per-iteration
    statistic is obtained from the shared weighted-covariance kernel (x_psi = V w / (w^H V w)); the stopping
    rule ``max_f ||delta_f|| < tol`` (ive.py:238-241) is evaluated every iteration ON THE DEVICE -- once it holds,
    the per-bin update becomes a no-op -- and read by the host every 20 epochs, so the result is the state of the
    epoch at which the reference stops, without a device-to-host round trip per epoch.
    """
    if getattr(X, "ndim", 0) != 3:
        raise ValueError("X must have shape (n_frames, n_freq, n_chan)")
    inp = _Input(X)
    T, F, M = inp.dev.shape
    lib = L.load()
    dev = inp.device
    mcode = _model_code(model, _OGIVE_MODELS)
    with torch.cuda.device(dev):
        st = _stream_ptr(dev)
        code = L.C64 if inp.dtype == torch.complex64 else L.C128
        plan = DemixPlan(1, T, F, M, 1, mcode, inp.dtype, dev)
        plan.load(inp.dev[None])
        status = lib.oiva_plan_status_ptr(plan.h)
        Cx = plan.cov  # (1, F, M, M)
        c128 = dict(dtype=torch.complex128, device=dev)
        Cinv = torch.empty((F, M, M), **c128)
        cnorm = torch.empty((F,), dtype=torch.float64, device=dev)
        L.check(lib.oiva_ogive_setup(_ptr(Cx), _ptr(Cinv), _ptr(cnorm), status, F, M, st), "oiva_ogive_setup")
        w = torch.zeros((F, M), **c128)
        a = torch.zeros((F, M), **c128)
        lam = torch.zeros((F,), dtype=torch.float64, device=dev)
        if W0 is not None:  # ive.py:129-130: w[:, :] = W0 with w of shape (F, M, 1)
            w.copy_(_prepare_W0(W0, 1, F, M, 1, dev)[0, :, :, 0])
        elif init_eig:  # ive.py:110-123: the un-conjugated leading eigenvector of np.linalg.eig
            evals = torch.empty((F, M), dtype=torch.float64, device=dev)
            evecs = torch.empty((F, M, M), **c128)
            L.check(lib.oiva_eigh(_ptr(Cx), _ptr(evals), _ptr(evecs), status, F, F, M, 1, st), "oiva_eigh")
            w.copy_(evecs[:, :, M - 1])
        else:
            w[:, 0] = 1.0  # ive.py:125-127
        L.check(lib.oiva_ogive_a_from_w(_ptr(w), _ptr(a), _ptr(Cx), F, M, st), "oiva_ogive_a_from_w")  # ive.py:168
        do_a = torch.full((F,), 1 if update == "mix" else 0, dtype=torch.uint8, device=dev)  # ive.py:170-175
        nch = lib.oiva_bin_groups(F)
        Tp = plan.Tp
        r2part = torch.empty((nch, Tp), dtype=torch.float64, device=dev)
        phi = torch.empty((Tp,), dtype=torch.float64, device=dev)
        Vg = torch.empty(lib.oiva_grouped_cov_bytes(1, F, M, 1), dtype=torch.uint8, device=dev)
        cov_ws_bytes = lib.oiva_weighted_cov_scratch_bytes(1, T, F, M, 1)
        cov_ws = torch.empty(max(cov_ws_bytes, 16), dtype=torch.uint8, device=dev)
        V = torch.empty((F, M, M), **c128)
        n_iter = int(n_iter)
        dhist = torch.zeros((max(n_iter, 1),), dtype=torch.float64, device=dev)  # max_f ||delta_f|| per epoch
        SYNC_EVERY = 20  # divides the callback period (100): convergence is known before every callback

        def project(wcur):
            Weff = torch.empty((F, M, 1), **c128)
            L.check(lib.oiva_projback_filters(_ptr(wcur), 1, _ptr(Cx), _ptr(Weff), F, M, 1, int(bool(proj_back)), st),
                    "oiva_projback_filters")
            Y = torch.empty((1, T, F, 1), dtype=inp.dtype, device=dev)
            L.check(lib.oiva_demix_output(plan.samples_ptr, _ptr(Weff), _ptr(Y), 1, T, F, M, 1, code, st),
                    "oiva_demix_output")
            return Y[0]

        def converged(upto):  # ive.py:238-241 for the epochs run so far (one small D2H copy)
            return upto > 0 and float(dhist[upto - 1].item()) < tol

        # the stopping rule is evaluated on the device every epoch (oiva_ogive_update_gated freezes the state at the epoch
        # the reference breaks at); the host only looks every SYNC_EVERY epochs.  Epochs run in blocks of 10 (one library
        # call each: oiva_ogive_iterate): the switching criterion (ive.py:187-188, every 10th epoch), the callback
        # (ive.py:194-200, every 100th) and the convergence check all fall on block boundaries.
        BLOCK = 10
        epoch = 0
        while epoch < n_iter:
            if epoch % SYNC_EVERY == 0 and converged(epoch):
                break
            if update == "switching":  # (epoch % 10 == 0 here)
                L.check(lib.oiva_ogive_switching(_ptr(a), _ptr(Cx), _ptr(cnorm), _ptr(do_a), F, M, st),
                        "oiva_ogive_switching")
            if callback is not None and epoch % 100 == 0:
                callback(inp.give_back(project(w)))
            n_blk = min(BLOCK, n_iter - epoch)
            L.check(lib.oiva_ogive_iterate(plan.samples_ptr, _ptr(w), _ptr(a), _ptr(lam), _ptr(r2part), _ptr(phi), _ptr(Vg),
                                           _ptr(cov_ws), cov_ws_bytes, _ptr(V), _ptr(Cx), _ptr(Cinv), _ptr(do_a),
                                           float(step_size), _ptr(dhist), epoch, n_blk, float(tol), T, F, M, mcode, code, st),
                    "oiva_ogive_iterate")
            epoch += n_blk
        Y = project(w)
        plan.raise_on_failure()
        Yo = inp.give_back(Y)
        if return_filters:
            return Yo, inp.give_back(w[:, :, None].contiguous(), inp.dtype)
        return Yo
