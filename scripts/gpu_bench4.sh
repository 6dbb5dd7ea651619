#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 4 gpurun_out/$name.log | cut -c1-3500 | tee -a gpurun_out/summary.txt
}
run bench512 900 python bench.py
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file gpurun_out/launches_b512.csv python bench.py --steps 2 --no-cpu
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_ip_update_tpb|k_demix_power" -s 190 -c 3 -o gpurun_out/prof_b512 python bench.py --steps 1 --no-cpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run refarm 600 python bench.py --impl reference --steps 2 --warmup 1
