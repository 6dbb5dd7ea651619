#!/bin/bash
# First-contact GPU check: run each group in its own process under `timeout`, so a hung kernel costs one
# group, not the whole call.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 15 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
PT="python -m pytest -q -x --timeout 240 -p no:cacheprovider"
run k_basic   400 $PT tests/test_kernels_gpu.py -k "relayout or demix_power or source_model or eigh or init_demix or ip_update or final_demix"
run k_cov_ld  400 $PT tests/test_kernels_gpu.py -k "weighted_covariance and True"
run k_cov_tma 400 $PT tests/test_kernels_gpu.py -k "weighted_covariance and not True"
run api       900 $PT tests/test_api_gpu.py
run smoke     200 python -c "import __graft_entry__ as g; g.smoke()"
