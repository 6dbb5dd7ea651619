#!/bin/bash
# Round-2 call P (after reverting the sweep-on-partials path): sweep on per-split partial covariances, single-launch source model, register bounds of the streaming
# kernels, 4-part covariance at M = K = 6; tests, all configs, determined batches, the bench.
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
TAILN=30 run r02p_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
TAILN=8 run r02p_resident 300 python scripts/check_resident.py
run r02p_kernels 600 python scripts/profile_configs.py cfg3,cfg4_b64,det4_b256,det6_b256,det8_b256
run r02p_configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3,cfg5
run r02p_bench 900 python bench.py --no-cpu --no-cfg5
