#!/bin/bash
# Round-2 final measurement set (one GPU): tests, smoke, the full bench line + reference arm, per-kernel breakdown of all
# shapes, configs, audio path, next rows (sweep, monitor, ILRMA), ncu launch list and full captures (raw CSV only).
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
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
run fin_pytest 1200 python -m pytest tests -q -m gpu --timeout 300
run fin_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run fin_bench_n1 1500 python bench.py
run fin_refarm 600 python bench.py --impl reference --steps 1 --warmup 1
TAILN=12 run fin_kernels 600 python scripts/profile_configs.py cfg1,cfg2,cfg3,cfg5,cfg5_shard8,cfg4_b64,det4_b256,det6_b256,det8_b256
run fin_configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3,cfg5 --cpu
run fin_audio 600 python scripts/bench_audio.py
run fin_nextrows 900 python scripts/bench_next_rows.py
run fin_ilrma 600 python scripts/bench_ilrma.py
run fin_ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5
run fin_ncu_bench 900 ncu --set full --clock-control none -k regex:"k_cov_sweep|k_demix_power" -s 60 -c 2 -o gpurun_out/fin_bench python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5
run fin_ncu_cfg5 900 ncu --set full --clock-control none -k regex:"k_cov_tiled|k_demix_staged" -s 4 -c 2 -o gpurun_out/fin_cfg5 python scripts/profile_configs.py cfg5
run fin_ncu_res 600 ncu --set full --clock-control none -k regex:"k_loop_resident" -s 2 -c 1 -o gpurun_out/fin_res python scripts/bench_configs.py --configs cfg1 --reps 2
for n in bench cfg5 res; do ncu -i gpurun_out/fin_$n.ncu-rep --page raw --csv > gpurun_out/fin_${n}_raw.csv 2>/dev/null; rm -f gpurun_out/fin_$n.ncu-rep; done
