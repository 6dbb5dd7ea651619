#!/bin/bash
# compute-sanitizer over the resident loop after the cluster / tagged-word / DSMEM changes (small shapes).
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
K1="resident_loop_synchronisation_modes_agree and 1500"
K2="resident_loop_equals_kernel_loop and (2-2-laplace or 3-1-gauss)"
run mem_res 900 compute-sanitizer --tool memcheck --error-exitcode 1 --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "($K1) or ($K2) or singular_mixture_does_not_stall"
run sync_res 900 compute-sanitizer --tool synccheck --error-exitcode 1 --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "($K1) or ($K2)"
run race_res 900 compute-sanitizer --tool racecheck --racecheck-report analysis --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "($K1) or ($K2)"
