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


def _worker(rank, world, port, out_dir, model):
    import torch.distributed as dist

    from overiva_b200.distributed import overiva_freq_sharded, shard_bins

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X = _mixture()
        F = X.shape[1]
        f0, f1 = shard_bins(F, world, rank)
        Y, W = overiva_freq_sharded(np.ascontiguousarray(X[:, f0:f1]), F, n_src=2, n_iter=10, model=model,
                                    return_filters=True)
        np.save(os.path.join(out_dir, "y_%d.npy" % rank), Y)
        np.save(os.path.join(out_dir, "w_%d.npy" % rank), W)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model", ["laplace", "gauss"])
def test_freq_sharded_two_gpus_nccl(model, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), model), nprocs=2, join=True)
    X = _mixture()
    Yo, Wo = orc.overiva(X, n_src=2, n_iter=10, model=model, return_filters=True)
    Y = np.concatenate([np.load(os.path.join(str(tmp_path), "y_%d.npy" % r)) for r in range(2)], axis=1)
    W = np.concatenate([np.load(os.path.join(str(tmp_path), "w_%d.npy" % r)) for r in range(2)], axis=0)
    assert rel_err(Y, Yo) < 1e-10 and rel_err(W, Wo) < 1e-10
