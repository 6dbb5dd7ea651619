#!/bin/bash
# Round-2 call J: fused covariance + sweep kernel (cov_sweep.cuh): parity against the two-kernel path, bench A/B.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-5000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
TAILN=30 run r02j_fused_test 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "fused_cov_sweep or ip_update"
run r02j_fullsize 600 python -m pytest tests/test_edge_gpu.py tests/test_api_gpu.py -q -m gpu -x --timeout 300
run r02j_bench_fused 900 python bench.py --no-cpu --no-e2e --no-configs --no-cfg5
run r02j_bench_two 900 env OIVA_NO_COV_SWEEP=1 python bench.py --no-cpu --no-e2e --no-configs --no-cfg5
run r02j_pytest 1200 python -m pytest tests -q -m gpu -x --timeout 300
