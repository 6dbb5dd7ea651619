"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: bin / batch sharding and the per-iteration
all-reduce of the source-model statistic.  The collective driver (overiva_b200.distributed.iterate_freq_sharded)
is the product code; the per-shard engine is a numpy stand-in built from the oracle's step functions, so the
test checks exactly what the N > 1 path adds: that sharding by bins + one K x T sum per iteration reproduces
the unsharded algorithm."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, rel_err
from oracle import overiva_oracle as orc
from overiva_b200.distributed import iterate_freq_sharded, shard_batch, shard_bins


def test_shard_bins_covers_everything_on_group_boundaries():
    for F in (1, 31, 32, 33, 257, 2049):
        for world in (1, 2, 3, 8):
            edges = [shard_bins(F, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == F
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            for f0, f1 in edges[:-1]:
                assert f1 % 32 == 0 or f1 == F
    assert shard_bins(2049, 8, 0) == (0, 256) and shard_bins(2049, 8, 7) == (1792, 2049)


def test_shard_batch():
    cuts = [shard_batch(4096, 8, r) for r in range(8)]
    assert cuts[0] == (0, 512) and cuts[-1] == (3584, 4096)
    assert sum(b - a for a, b in cuts) == 4096


class NumpyEngine:
    """Stand-in for the CUDA plan on one shard of bins (test-only)."""

    def __init__(self, X_local, K, model, F_total):
        self.K, self.model, self.F_total = K, model, F_total
        self.C = orc.input_covariance(X_local)
        self.What = orc.init_demixing(self.C, K)
        self.Xf = np.ascontiguousarray(X_local.swapaxes(0, 1))
        self.r2 = torch.zeros((1, K, X_local.shape[0]), dtype=torch.float64)

    def power(self):
        r2 = orc.demix_power(self.Xf, self.What[:, :, : self.K])  # (T, K)
        self.r2[0] = torch.from_numpy(np.ascontiguousarray(r2.T))
        return self.r2

    def update(self):
        orc.iterate_once(self.Xf, self.What, self.C, self.K, self.model, n_freq_total=self.F_total,
                         r2=self.r2[0].numpy().T.copy())


def _worker(rank, world, port, name, model, n_iter, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import overiva_b200.distributed as D

    D.GROUP = 4  # the golden inputs have < 32 bins: cut them with a small group size so that BOTH ranks own bins
    try:
        case = load_golden(name)
        X = case["X"]
        K = case["kwargs"].get("n_src") or X.shape[2]
        F = X.shape[1]
        f0, f1 = D.shard_bins(F, world, rank)
        assert f1 > f0, "every rank must own bins in this test"
        eng = NumpyEngine(X[:, f0:f1], K, model, F)
        iterate_freq_sharded(eng, n_iter)
        np.save(os.path.join(out_dir, "what_%d.npy" % rank), eng.What)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("name,model", [("overiva_m2", "laplace"), ("overiva_gauss_eig_m8k2", "gauss")])
def test_frequency_sharded_loop_world2_gloo(name, model, tmp_path):
    world, n_iter = 2, 6
    mp.spawn(_worker, args=(world, _free_port(), name, model, n_iter, str(tmp_path)), nprocs=world, join=True)
    case = load_golden(name)
    X = case["X"]
    K = case["kwargs"].get("n_src") or X.shape[2]
    C = orc.input_covariance(X)
    full = orc.init_demixing(C, K)
    Xf = np.ascontiguousarray(X.swapaxes(0, 1))
    for _ in range(6):
        orc.iterate_once(Xf, full, C, K, model)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "what_%d.npy" % r)) for r in range(2)], axis=0)
    assert got.shape == full.shape
    assert rel_err(got, full) < 1e-12
