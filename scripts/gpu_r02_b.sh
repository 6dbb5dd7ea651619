#!/bin/bash
# Round-2 call B: tiled covariance v2 (cluster multicast), staged output kernel, the full new bench line.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | cut -c1-6000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
run r02b_tiled 240 python scripts/check_tiled.py
if ! grep -q TILED_OK gpurun_out/r02b_tiled.log; then export OIVA_COV_NO_TILED=1; echo "TILED KERNEL DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02b_pytest 900 python -m pytest tests -q -m gpu -x
run r02b_kernels 600 python scripts/profile_configs.py cfg3,cfg5,cfg5_shard8
run r02b_bench_n1 1500 python bench.py
run r02b_ncu_cfg5 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov_tiled|k_demix_staged" -s 4 -c 2 -o gpurun_out/r02b_cfg5 python scripts/profile_configs.py cfg5
ncu -i gpurun_out/r02b_cfg5.ncu-rep --page raw --csv > gpurun_out/r02b_cfg5_raw.csv 2>/dev/null
