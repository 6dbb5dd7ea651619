#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.log | cut -c1-1800 | tee -a gpurun_out/summary.txt
}
run pytest_all 900 python -m pytest tests -x -q -m gpu
run bench_fused 600 python bench.py --steps 5 --no-cpu --no-e2e
OIVA_NO_RELAYOUT_COV=1 run bench_unfused 600 python bench.py --steps 5 --no-cpu --no-e2e
