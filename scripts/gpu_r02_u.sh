#!/bin/bash
# Round-2 call U: whole-chunk path only for M <= 6, four sources per warp in the streaming kernels at M = 7, 8.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
TAILN=30 run r02u_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02u_kernels 600 python scripts/profile_configs.py cfg3,det6_b256,det8_b256
run r02u_ilrma 600 python scripts/bench_ilrma.py
run r02u_bench 900 python bench.py --no-cpu --no-e2e --no-cfg5
