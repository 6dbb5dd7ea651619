#!/bin/bash
# Round-2 call K: grouped init / single-launch output / lazy full C, ramped chunk schedule of the host pipelines,
# ring shape of the fused covariance + sweep kernel, ncu capture of the resident loop at cfg2.
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
TAILN=30 run r02k_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02k_bench 1500 python bench.py --no-cfg5
run r02k_bench_tc2 600 env OIVA_SWEEP_TC=2 python bench.py --no-cpu --no-e2e --no-configs --no-cfg5
run r02k_bench_tc2s3 600 env OIVA_SWEEP_TC=2 OIVA_SWEEP_STAGES=3 python bench.py --no-cpu --no-e2e --no-configs --no-cfg5
run r02k_bench_tc4s3 600 env OIVA_SWEEP_STAGES=3 python bench.py --no-cpu --no-e2e --no-configs --no-cfg5
run r02k_ncu_res2 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02k_res2 python scripts/bench_configs.py --configs cfg2 --reps 2
ncu -i gpurun_out/r02k_res2.ncu-rep --page source --csv > gpurun_out/r02k_res2_sass.csv 2>/dev/null
ncu -i gpurun_out/r02k_res2.ncu-rep --page raw --csv > gpurun_out/r02k_res2_raw.csv 2>/dev/null
