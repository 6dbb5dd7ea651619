#!/bin/bash
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
TAILN=5 run r02g_resident 240 python scripts/check_resident.py
run r02g_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02g_kernels 300 python scripts/profile_configs.py cfg5
run r02g_ncu_res1 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02g_res1 python scripts/bench_configs.py --configs cfg1 --reps 2
run r02g_ncu_res2 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02g_res2 python scripts/bench_configs.py --configs cfg2 --reps 2
