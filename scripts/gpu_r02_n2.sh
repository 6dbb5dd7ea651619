#!/bin/bash
# 2-GPU call: NCCL parity tests of the frequency-sharded loop and the full bench line at N = 2 (rank-safe legs).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-7000 | tee -a gpurun_out/summary.txt
  tail -n 8 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
nvidia-smi topo -m > gpurun_out/r02_n2_topo.txt 2>&1; free -g >> gpurun_out/r02_n2_topo.txt; nproc >> gpurun_out/r02_n2_topo.txt
run r02_n2_dist 600 python -m pytest tests/test_distributed_gpu.py -q -m gpu --timeout 300
run r02_n2_bench 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
run r02_n2_ref 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1
