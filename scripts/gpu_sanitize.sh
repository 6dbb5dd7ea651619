#!/bin/bash
# compute-sanitizer (memcheck) over small GPU tests of the kernels added in round 2: two-lanes-per-bin sweep, grouped init /
# output, ILRMA (per-bin weighted covariance, NMF kernels), cross-correlations, resident loop, frame-parallel source model.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Race" gpurun_out/$name.log | tail -n 12 | cut -c1-300 | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.err | cut -c1-600 | tee -a gpurun_out/summary.txt
}
SAN="compute-sanitizer --tool memcheck --error-exitcode 1 --target-processes all"
run san_sweep 900 $SAN python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "(ip_update_sweep and (7-7 or 8-8 or 6-6 or 4-2)) or init_demix_grouped or final_demix or source_model"
run san_ilrma 900 $SAN python -m pytest tests/test_ilrma.py -q -m gpu -x -k "matches_the_oracle and (2-2 or 3-2 or 8-2)"
run san_metrics 900 $SAN python -m pytest tests/test_metrics.py -q -m gpu -x -k "32-5000 or 1-1000"
run san_api 900 $SAN python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "golden and (auxiva_gauss_m4 or auxiva_laplace_m8 or auxiva_pca_laplace_m5k2 or ogive_demix_laplace_m4)"
