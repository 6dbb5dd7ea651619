#!/bin/bash
# compute-sanitizer racecheck / synccheck over small tests of the shared-memory kernels (resident loop, covariance rings,
# staged statistic kernel, two-lanes-per-bin sweep).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/$name.log | sort | uniq -c | sort -rn | head -n 8 | cut -c1-300 | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.err | cut -c1-600 | tee -a gpurun_out/summary.txt
}
T1='tests/test_api_gpu.py -q -m gpu -x -k "resident_loop_equals_kernel_loop and (4-2-laplace or 6-6-laplace)"'
run sync_res 900 compute-sanitizer --tool synccheck --error-exitcode 1 --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "resident_loop_equals_kernel_loop and (4-2-laplace or 6-6-laplace or 2-2-laplace)"
run race_res 900 compute-sanitizer --tool racecheck --racecheck-report analysis --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "resident_loop_equals_kernel_loop and (2-2-laplace or 3-1-gauss)"
run race_sweep 900 compute-sanitizer --tool racecheck --racecheck-report analysis --target-processes all python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "ip_update_sweep and (8-8 or 7-7 or 16-4)"
