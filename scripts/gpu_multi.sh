#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 4 gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt
}
nvidia-smi -L | tee -a gpurun_out/summary.txt
run dist_tests 600 python -m pytest -q --timeout 300 -p no:cacheprovider --tb=short tests/test_distributed_gpu.py
run bench_n1 600 python bench.py --no-cpu
run bench_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3
