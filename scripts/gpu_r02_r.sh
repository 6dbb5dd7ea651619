#!/bin/bash
# Round-2 call R: device cross-correlations / 512-tap metric, ILRMA in the sweep; the full suite.
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
TAILN=40 run r02r_metrics 600 python -m pytest tests/test_metrics.py tests/test_monitor.py tests/test_sweep.py tests/test_ilrma.py -q -m gpu --timeout 300
TAILN=30 run r02r_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
