#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 12 gpurun_out/$name.log | cut -c1-3500 | tee -a gpurun_out/summary.txt
}
run bench_n1 900 python bench.py
run chunks 600 python scripts/bench_chunks.py
OIVA_NO_GRAPH=1 run ncu_cfgs 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfgs.csv python scripts/bench_configs.py --configs cfg1,cfg2,cfg3 --reps 1
run ncu_stft 600 ncu --set full --clock-control none --import-source on -k regex:"k_stft_analysis|k_stft_frames" -c 2 -o gpurun_out/prof_stft python scripts/bench_audio.py --mixtures 128 --reps 1
