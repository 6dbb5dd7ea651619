#!/bin/bash
# compute-sanitizer over the tracked-inverse determined sweep of the resident loop (role-specialised warps, named barriers).
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
K1="tracked_inverse_sweep and (3-9000 or 5-12000)"
run mem_trk 900 compute-sanitizer --tool memcheck --error-exitcode 1 --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "$K1"
run sync_trk 900 compute-sanitizer --tool synccheck --error-exitcode 1 --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "$K1"
run race_trk 900 compute-sanitizer --tool racecheck --racecheck-report analysis --target-processes all python -m pytest tests/test_api_gpu.py -q -m gpu -x -k "tracked_inverse_sweep and 3-9000"
