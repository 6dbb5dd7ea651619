#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 600 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 4 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n rc=$?"; grep '^{' gpurun_out/bench_n$n.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e'].get('value'), 'frac', d['roofline']['frac'])"
done
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n scripts/bench_freq_sharded.py > gpurun_out/cfg5_n$n.log 2>&1; echo "cfg5 n=$n rc=$?"; grep '^{' gpurun_out/cfg5_n$n.log | cut -c1-400
done
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider --tb=short tests/test_distributed_gpu.py 2>&1 | tail -n 2
