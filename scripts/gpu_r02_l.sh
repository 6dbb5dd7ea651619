#!/bin/bash
# Round-2 call L: lighter grid barrier / handshakes of the resident loop, frame-parallel source model for one short
# mixture, projection-back scales in their own kernel for M >= 9; the full test-suite.
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
TAILN=8 run r02l_resident 300 python scripts/check_resident.py
TAILN=30 run r02l_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
run r02l_configs 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3,cfg5
run r02l_kernels 600 python scripts/profile_configs.py cfg3,cfg5
run r02l_bench 900 python bench.py --no-cpu --no-e2e --no-cfg5
