#!/bin/bash
# 4-GPU call: the full bench line at N = 4 (batch split + frequency-sharded cfg5 leg with its NCCL all-reduce),
# the NCCL parity tests, the reference arm as the driver launches it.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-9000 | tee -a gpurun_out/summary.txt
  tail -n 8 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
nvidia-smi topo -m > gpurun_out/n4_topo.txt 2>&1; free -g >> gpurun_out/n4_topo.txt; nproc >> gpurun_out/n4_topo.txt; lscpu | grep -i numa >> gpurun_out/n4_topo.txt
run n4_bench 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3
