#!/bin/bash
# Round-1b measurement set: tests, headline bench, reference arm, launch list + full capture of the top kernels,
# all BASELINE configs, audio path, next-row measurements.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | cut -c1-3500 | tee -a gpurun_out/summary.txt
}
run pytest_gpu 900 python -m pytest tests -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_n1 900 python bench.py
run refarm 600 python bench.py --impl reference --steps 2 --warmup 1
run configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3,cfg5 --cpu
run audio 600 python scripts/bench_audio.py
run nextrows 900 python scripts/bench_next_rows.py
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 261 -c 174 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --no-cpu --no-e2e
run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_ip_update_tpb|k_demix_power" -s 190 -c 3 -o gpurun_out/prof_r1b python bench.py --steps 1 --no-cpu --no-e2e
