#!/bin/bash
# Round-2 call S: two-lanes-per-bin determined sweep (M = 7, 8), covariance ring depth from free shared memory, ILRMA
# epochs in one library call; the full suite and the affected shapes.
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
TAILN=30 run r02s_sweep 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "ip_update or fused"
TAILN=30 run r02s_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02s_kernels 600 python scripts/profile_configs.py cfg3,cfg4_b64,det6_b256,det8_b256
run r02s_configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3
run r02s_ilrma 600 python scripts/bench_ilrma.py
