#!/bin/bash
# Round-2 call E: resident loop v3, tiled covariance v3 (out-of-line producer, mid-chunk polling); ncu of cfg2 resident.
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
TAILN=5 run r02e_resident 240 python scripts/check_resident.py
if ! grep -q RESIDENT_OK gpurun_out/r02e_resident.log; then export OIVA_NO_RESIDENT=1; echo "RESIDENT LOOP DISABLED" | tee -a gpurun_out/summary.txt; fi
TAILN=2 run r02e_tiled 240 python scripts/check_tiled.py
if ! grep -q TILED_OK gpurun_out/r02e_tiled.log; then export OIVA_COV_NO_TILED=1; echo "TILED KERNEL DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02e_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02e_kernels 600 python scripts/profile_configs.py cfg5,cfg5_shard8
run r02e_ncu_res2 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02e_res2 python scripts/bench_configs.py --configs cfg2 --reps 2
run r02e_ncu_cfg5 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov_tiled" -s 2 -c 1 -o gpurun_out/r02e_cfg5 python scripts/profile_configs.py cfg5
ncu -i gpurun_out/r02e_cfg5.ncu-rep --page raw --csv > gpurun_out/r02e_cfg5_raw.csv 2>/dev/null
