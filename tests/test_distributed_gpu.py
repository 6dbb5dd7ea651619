"""GPU tests of the multi-GPU drivers.  The single-process case runs everywhere; the NCCL case needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_distributed_gpu.py -m gpu`) and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from conftest import rel_err
from oracle import overiva_oracle as orc
from overiva_b200.synth import small_test_mixture

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _mixture():
    return small_test_mixture(400, 4, 2, n_samples=6000, frame=256, hop=128)  # F = 129 bins: 5 lane groups


def test_freq_sharded_single_rank_equals_overiva():
    import overiva_b200 as ob
    from overiva_b200.distributed import overiva_freq_sharded

    X = _mixture()
    Y1, W1 = ob.overiva(X, n_src=2, n_iter=8, return_filters=True)
    Y2, W2 = overiva_freq_sharded(X, X.shape[1], n_src=2, n_iter=8, return_filters=True)
    assert rel_err(Y2, Y1) < 1e-12 and rel_err(W2, W1) < 1e-12


def _mixture16():
    # BASELINE config 5's channel / source counts (M = 16, K = 4) on a mixture long enough for frame splitting:
    # T = 2082 frames, F = 97 bins (4 lane groups, the last one ragged).  4 targets + 16 interferers >= 16 channels keeps
    # the reference well conditioned (a 1e-15 perturbation of X moves ITS W by 2e-13 after 10 iterations; with fewer
    # sources than channels it is 1e-11 and a 1e-10 bound would test the input, not the implementation)
    return small_test_mixture(401, 16, 4, n_samples=200000, frame=192, hop=96, n_interferers=16)


@pytest.mark.parametrize("model", ["laplace", "gauss"])
def test_freq_sharded_m16k4_single_rank_vs_oracle(model):
    """World size 1 still goes through iterate_freq_sharded (power -> [all-reduce] -> update per epoch) and the
    many-channel kernels (tiled covariance with frame splits, staged demix, row-owner sweep): vs the oracle."""
    from overiva_b200.distributed import overiva_freq_sharded

    X = _mixture16()
    assert X.shape[0] >= 2000 and X.shape[1] >= 96
    Yo, Wo = orc.overiva(X, n_src=4, n_iter=8, model=model, return_filters=True)
    Y, W = overiva_freq_sharded(X, X.shape[1], n_src=4, n_iter=8, model=model, return_filters=True)
    assert rel_err(Y, Yo) < 1e-10 and rel_err(W, Wo) < 1e-10


@pytest.mark.parametrize("model,n_shards", [("laplace", 3), ("gauss", 2)])
def test_freq_sharded_emulated_ranks_m16k4_vs_oracle(model, n_shards):
    """The arithmetic of the N-rank run on ONE GPU: one plan per bin shard, the source-model statistic summed over
    the shards between power() and update() exactly as the all-reduce does (what the 1-GPU test box can check of
    the multi-rank path; the NCCL transport itself is test_freq_sharded_two_gpus_nccl)."""
    from overiva_b200 import _lib as L
    from overiva_b200.core import DemixPlan
    from overiva_b200.distributed import shard_bins

    X = _mixture16()
    T, F, M = X.shape
    K, n_iter = 4, 8
    code = L.MODEL_LAPLACE if model == "laplace" else L.MODEL_GAUSS
    dev = torch.device("cuda", 0)
    plans, edges = [], []
    for r in range(n_shards):
        f0, f1 = shard_bins(F, n_shards, r)
        edges.append((f0, f1))
        plan = DemixPlan(1, T, f1 - f0, M, K, code, torch.complex128, dev, n_freq_total=F)
        plan.load(torch.from_numpy(np.ascontiguousarray(X[None, :, f0:f1])).to(dev))
        plan.init(L.INIT_EYE)
        plans.append(plan)
    for _ in range(n_iter):
        for plan in plans:
            plan.power()
        total = sum(plan.r2 for plan in plans)
        for plan in plans:
            plan.r2.copy_(total)
            plan.update()
    Y = np.concatenate([plan.output(True)[0].cpu().numpy() for plan in plans], axis=1)
    W = np.concatenate([plan.filters()[0].cpu().numpy() for plan in plans], axis=0)
    for plan in plans:
        plan.raise_on_failure()
    Yo, Wo = orc.overiva(X, n_src=K, n_iter=n_iter, model=model, return_filters=True)
    assert rel_err(Y, Yo) < 1e-10 and rel_err(W, Wo) < 1e-10


def _worker(rank, world, port, out_dir, model, big=False):
    import torch.distributed as dist

    from overiva_b200.distributed import overiva_freq_sharded, shard_bins

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X = _mixture16() if big else _mixture()
        F = X.shape[1]
        f0, f1 = shard_bins(F, world, rank)
        Y, W = overiva_freq_sharded(np.ascontiguousarray(X[:, f0:f1]), F, n_src=4 if big else 2, n_iter=10, model=model,
                                    return_filters=True)
        np.save(os.path.join(out_dir, "y_%d.npy" % rank), Y)
        np.save(os.path.join(out_dir, "w_%d.npy" % rank), W)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,big", [("laplace", False), ("gauss", False), ("laplace", True), ("gauss", True)])
def test_freq_sharded_two_gpus_nccl(model, big, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), model, big), nprocs=2, join=True)
    X = _mixture16() if big else _mixture()
    Yo, Wo = orc.overiva(X, n_src=4 if big else 2, n_iter=10, model=model, return_filters=True)
    Y = np.concatenate([np.load(os.path.join(str(tmp_path), "y_%d.npy" % r)) for r in range(2)], axis=1)
    W = np.concatenate([np.load(os.path.join(str(tmp_path), "w_%d.npy" % r)) for r in range(2)], axis=0)
    assert rel_err(Y, Yo) < 1e-10 and rel_err(W, Wo) < 1e-10
