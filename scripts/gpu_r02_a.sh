#!/bin/bash
# Round-2 call A: tests after the library clean-up, the new bench line, fp64 DFMA/DMMA rates, per-kernel breakdown of
# every BASELINE shape, and the ncu evidence for the many-channel covariance kernel (cfg5).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | cut -c1-6000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
run r02a_tiled 240 python scripts/check_tiled.py
if ! grep -q TILED_OK gpurun_out/r02a_tiled.log; then export OIVA_COV_NO_TILED=1; echo "TILED KERNEL DISABLED" | tee -a gpurun_out/summary.txt; fi
run r02a_pytest 900 python -m pytest tests -q -m gpu -x
run r02a_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run r02a_bench_n1 1200 python bench.py
run r02a_kernels 600 python scripts/profile_configs.py cfg1,cfg2,cfg3,cfg5,cfg5_shard8
run r02a_ncu_cfg5 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov_blocked|k_cov_tiled|k_demix_power" -s 4 -c 3 -o gpurun_out/r02a_cfg5 python scripts/profile_configs.py cfg5
run r02a_kernels_blocked 600 env OIVA_COV_NO_TILED=1 python scripts/profile_configs.py cfg5,cfg5_shard8
ncu -i gpurun_out/r02a_cfg5.ncu-rep --page raw --csv > gpurun_out/r02a_cfg5_raw.csv 2>/dev/null
