#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | cut -c1-700 | tee -a gpurun_out/summary.txt
}
run pytest_stft 600 python -m pytest tests/test_stft.py tests/test_monitor.py tests/test_sweep.py -x -q -m gpu
run audio2 600 python scripts/bench_audio.py --reps 2
