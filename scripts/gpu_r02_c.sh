#!/bin/bash
# Round-2 call C: the persistent single-launch loop, staged demix kernels with sub-block warps, tiled covariance v2b.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-6000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
TAILN=24 run r02c_resident 240 python scripts/check_resident.py
if ! grep -q RESIDENT_OK gpurun_out/r02c_resident.log; then export OIVA_NO_RESIDENT=1; echo "RESIDENT LOOP DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02c_tiled 240 python scripts/check_tiled.py
if ! grep -q TILED_OK gpurun_out/r02c_tiled.log; then export OIVA_COV_NO_TILED=1; echo "TILED KERNEL DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02c_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02c_kernels 600 python scripts/profile_configs.py cfg3,cfg5,cfg5_shard8,cfg4_b64
run r02c_bench_n1 1500 python bench.py
