#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 12 gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt
}
run pytest_new 600 python -m pytest tests/test_stft.py tests/test_sweep.py tests/test_monitor.py -q -m gpu
run audio 600 python scripts/bench_audio.py
