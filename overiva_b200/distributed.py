"""Multi-GPU drivers (one process per GPU, ``torch.distributed``; NCCL over NVLink on the 8xB200 box).

The path shards in two natural ways:

* **many independent mixtures** (the role of the reference's task farm ``rrtools`` + ``overiva_sim.one_loop``):
  :func:`overiva_batch_sharded` -- every rank separates its own slice of the batch with
  :func:`overiva_b200.overiva_batch`; there is NO collective in the data path.
* **one long mixture**: :func:`overiva_freq_sharded` -- every rank holds a contiguous slice of the
  frequency bins.  Covariances, demixing matrices, the IP sweep, the eigen-initialisation, the final demix
  and the projection back are all per-bin; the only cross-bin quantity of the algorithm is the source-model
  statistic ``r2[k, t] = sum_f |y_k(f, t)|^2`` (``overiva.py:152-155``), so the loop needs exactly one
  sum-all-reduce of ``K x T`` doubles per iteration (squared partial sums are reduced, the square root comes
  after; the gauss model divides by the FULL number of bins).

:func:`iterate_freq_sharded` is written against a tiny engine protocol (``power() -> tensor``,
``update()``) so that the collective logic can be exercised on CPU with the gloo backend and a stand-in
engine (``tests/test_distributed_cpu.py``); the product engine is :class:`overiva_b200.core.DemixPlan`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib as L
from .core import DemixPlan, _Input, _model_code, _prepare_W0, overiva_batch

GROUP = 32  # bins per lane group of the CUDA kernels: shard boundaries are aligned to it


def shard_bins(n_freq: int, world_size: int, rank: int):
    """Contiguous split of ``n_freq`` bins over ``world_size`` ranks, boundaries on multiples of 32 bins
    (so no lane group straddles two ranks).  Returns ``(f_begin, f_end)``; a rank may get an empty range
    when there are fewer groups than ranks."""
    n_groups = (n_freq + GROUP - 1) // GROUP
    g0 = n_groups * rank // world_size
    g1 = n_groups * (rank + 1) // world_size
    return min(g0 * GROUP, n_freq), min(g1 * GROUP, n_freq)


def shard_batch(n_batch: int, world_size: int, rank: int):
    """Contiguous split of a batch of independent mixtures."""
    return n_batch * rank // world_size, n_batch * (rank + 1) // world_size


def iterate_freq_sharded(engine, n_iter: int, group=None):
    """``n_iter`` epochs of the loop on a frequency shard: local partial statistic -> sum all-reduce ->
    local source model + weighted covariance + IP sweep (``overiva.py:138-190``)."""
    for _ in range(int(n_iter)):
        r2 = engine.power()  # (B, K, Tp) view of the engine's statistic buffer, summed over the LOCAL bins
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(r2, op=dist.ReduceOp.SUM, group=group)
        engine.update()  # reads the (now global) statistic from the same buffer


class _PlanEngine:
    """Engine protocol over a CUDA :class:`DemixPlan`."""

    def __init__(self, plan: DemixPlan):
        self.plan = plan
        self._r2 = plan.r2

    def power(self):
        self.plan.power()
        return self._r2

    def update(self):
        self.plan.update()


class _ZeroEngine:
    """A rank that owns no bins still takes part in every all-reduce."""

    def __init__(self, zeros):
        self.zeros = zeros

    def power(self):
        return self.zeros.zero_()

    def update(self):
        pass


def overiva_freq_sharded(X_local, n_freq_total, n_src=None, n_iter=20, proj_back=True, W0=None, model="laplace",
                         init_eig=False, return_filters=False, group=None):
    """OverIVA on ONE mixture whose bins are sharded across the ranks of ``group``.

    ``X_local``: this rank's ``(n_frames, n_freq_local, n_chan)`` slice (see :func:`shard_bins`);
    ``n_freq_total``: the F of the whole mixture.  ``W0`` (if given) is the local ``(n_freq_local, n_chan,
    n_src)`` slice.  Returns this rank's ``Y (n_frames, n_freq_local, n_src)`` [and ``W``]; concatenating the
    ranks' outputs along the frequency axis gives exactly what :func:`overiva_b200.overiva` returns for the
    full mixture (up to the summation order of the all-reduce).
    """
    inp = _Input(X_local)
    T, F, M = inp.dev.shape
    K = M if n_src is None else int(n_src)
    if not (1 <= K <= M):
        raise ValueError("n_src=%d must be in 1..n_chan=%d" % (K, M))
    if F == 0:  # more ranks than bin groups: contribute zeros to every all-reduce, return empty slices
        Tp = L.load().oiva_frame_pitch(T)
        iterate_freq_sharded(_ZeroEngine(torch.zeros((1, K, Tp), dtype=torch.float64, device=inp.device)), n_iter,
                             group)
        Yo = inp.give_back(torch.empty((T, 0, K), dtype=inp.dtype, device=inp.device))
        if return_filters:
            return Yo, inp.give_back(torch.empty((0, M, K), dtype=inp.dtype, device=inp.device))
        return Yo
    with torch.cuda.device(inp.device):
        plan = DemixPlan(1, T, F, M, K, _model_code(model), inp.dtype, inp.device, n_freq_total=int(n_freq_total))
        plan.load(inp.dev[None])
        if W0 is not None:
            plan.init(L.INIT_W0, _prepare_W0(W0, 1, F, M, K, inp.device))
        else:
            plan.init(L.INIT_EIG if init_eig else L.INIT_EYE)
        iterate_freq_sharded(_PlanEngine(plan), n_iter, group)
        Y = plan.output(proj_back)
        W = plan.filters() if return_filters else None
        plan.raise_on_failure()
        Yo = inp.give_back(Y[0])
        if return_filters:
            return Yo, inp.give_back(W[0], inp.dtype)
        return Yo


def overiva_batch_sharded(X_local, **kwargs):
    """This rank's share of a batch of independent mixtures (see :func:`shard_batch`): plain
    :func:`overiva_b200.overiva_batch`, no communication."""
    return overiva_batch(X_local, **kwargs)
