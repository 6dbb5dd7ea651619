#!/bin/bash
# Round-2 call N: tests after the projection-back / eigh changes, determined batches again (output kernel regression),
# ncu launch list of cfg3, small captures of the determined kernels (no source import: keep gpurun_out small).
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
TAILN=30 run r02n_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02n_kernels 600 python scripts/profile_configs.py cfg3,det4_b256,det6_b256,det8_b256
run r02n_configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3
run r02n_list3 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/r02n_cfg3_launches.csv python scripts/bench_configs.py --configs cfg3 --reps 2
run r02n_ncu6 900 ncu --set full --clock-control none -k regex:"k_cov|k_demix|k_ip_update" -s 8 -c 4 -o gpurun_out/r02n_det6 python scripts/profile_configs.py det6_b256
run r02n_ncu8 900 ncu --set full --clock-control none -k regex:"k_cov|k_demix|k_ip_update" -s 10 -c 5 -o gpurun_out/r02n_det8 python scripts/profile_configs.py det8_b256
for n in det6 det8; do ncu -i gpurun_out/r02n_$n.ncu-rep --page raw --csv > gpurun_out/r02n_${n}_raw.csv 2>/dev/null; rm -f gpurun_out/r02n_$n.ncu-rep; done
