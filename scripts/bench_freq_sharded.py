#!/usr/bin/env python
"""BASELINE config 5 (one 10-minute 48 kHz mixture, M=16, K=4, T=14061, F=2049) with the frequency bins sharded
across the ranks and one K x T all-reduce of the source-model statistic per epoch (NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_freq_sharded.py

Each rank draws its own bins (synthetic STFT-domain data), so the run measures the sharded loop, not parity (parity
of the sharded loop is tests/test_distributed_gpu.py).  Strong scaling: the mixture is fixed, N varies.
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from overiva_b200.distributed import overiva_freq_sharded, shard_bins  # noqa: E402
from overiva_b200.synth import stft_domain_batch_torch  # noqa: E402

T, F, M, K, N_ITER, SECS = 14061, 2049, 16, 4, 20, 600.0


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    f0, f1 = shard_bins(F, world, rank)
    X = stft_domain_batch_torch(1, T, f1 - f0, M, K, seed=500 + rank, device=dev, chunk=1)[0]
    times = []
    for it in range(2 + 5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Y = overiva_freq_sharded(X, F, n_src=K, n_iter=N_ITER, model="laplace")
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 2:
            times.append(float(t.item()))
    assert bool(torch.isfinite(Y.real).all())
    if rank == 0:
        ms = statistics.median(times)
        alg = (2 * N_ITER + 2) * F * T * M * 16 + F * T * K * 16
        print(json.dumps({
            "config": "cfg5 frequency-sharded", "n_gpus": world, "shape": {"T": T, "F": F, "M": M, "K": K},
            "bins_per_rank": f1 - f0, "ms_per_call": ms, "mixture_s_per_s": SECS / (ms / 1e3),
            "algorithmic_GBps_aggregate": alg / (ms / 1e3) / 1e9, "scaling": "strong",
            "collective": "all_reduce(SUM) of (K, Tp) float64 = %d bytes per epoch" % (K * ((T + 31) // 32 * 32) * 8),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
