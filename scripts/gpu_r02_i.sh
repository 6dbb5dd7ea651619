#!/bin/bash
# Round-2 evidence set: tests, smoke, the full bench line + reference arm, per-kernel breakdown of every BASELINE shape,
# ncu launch list of the bench step and full captures of the loop kernels (bench shape), the tiled covariance (cfg5)
# and the resident loop (cfg1).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-7000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
run r02i_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02i_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run r02i_bench_n1 1500 python bench.py
run r02i_refarm 600 python bench.py --impl reference --steps 1 --warmup 1
TAILN=12 run r02i_kernels 600 python scripts/profile_configs.py cfg1,cfg2,cfg3,cfg5,cfg5_shard8,cfg4_b64
run r02i_ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5
run r02i_ncu_bench 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_ip_update_tpb|k_demix_power" -s 100 -c 3 -o gpurun_out/r02i_bench python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5
run r02i_ncu_cfg5 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov_tiled|k_demix" -s 4 -c 2 -o gpurun_out/r02i_cfg5 python scripts/profile_configs.py cfg5
run r02i_ncu_res 600 ncu --set full --clock-control none --import-source on -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/r02i_res python scripts/bench_configs.py --configs cfg1 --reps 2
for n in bench cfg5 res; do ncu -i gpurun_out/r02i_$n.ncu-rep --page raw --csv > gpurun_out/r02i_${n}_raw.csv 2>/dev/null; done
