#!/bin/bash
# Round-2 call D: resident loop v2 (local statistic, row-owner determined sweep), plan_run, long-mixture source model;
# ncu captures of the tiled covariance kernel (cfg5) and the resident loop (cfg1).
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
TAILN=24 run r02d_resident 240 python scripts/check_resident.py
if ! grep -q RESIDENT_OK gpurun_out/r02d_resident.log; then export OIVA_NO_RESIDENT=1; echo "RESIDENT LOOP DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02d_tiled 240 python scripts/check_tiled.py
if ! grep -q TILED_OK gpurun_out/r02d_tiled.log; then export OIVA_COV_NO_TILED=1; echo "TILED KERNEL DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02d_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02d_kernels 600 python scripts/profile_configs.py cfg5,cfg5_shard8
run r02d_bench_n1 1500 python bench.py
run r02d_ncu_cfg5 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov_tiled" -s 2 -c 1 -o gpurun_out/r02d_cfg5 python scripts/profile_configs.py cfg5
run r02d_ncu_res 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02d_res python scripts/bench_configs.py --configs cfg1 --reps 2
ncu -i gpurun_out/r02d_cfg5.ncu-rep --page raw --csv > gpurun_out/r02d_cfg5_raw.csv 2>/dev/null
ncu -i gpurun_out/r02d_res.ncu-rep --page raw --csv > gpurun_out/r02d_res_raw.csv 2>/dev/null
