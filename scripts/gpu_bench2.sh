#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 8 gpurun_out/$name.log | cut -c1-2500 | tee -a gpurun_out/summary.txt
}
PT="python -m pytest -q --timeout 240 -p no:cacheprovider --tb=line"
run kernels 600 $PT tests/test_kernels_gpu.py
run api 900 $PT tests/test_api_gpu.py
run bench64 300 python bench.py --batch 64 --steps 3 --no-cpu
run bench512 900 python bench.py --steps 5 --no-cpu
run ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_b64.csv python bench.py --batch 64 --steps 1 --no-cpu
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_ip_update|k_demix_power" -s 12 -c 3 -o gpurun_out/prof_b64 python bench.py --batch 64 --steps 1 --no-cpu
