#!/bin/bash
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
TAILN=16 run r02h_resident 300 python scripts/check_resident.py
if ! grep -q RESIDENT_OK gpurun_out/r02h_resident.log; then export OIVA_NO_RESIDENT=1; echo "RESIDENT LOOP DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02h_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
