#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 240 -p no:cacheprovider --tb=short"
timeout 600 $PT tests/test_kernels_gpu.py > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?"; tail -n 4 gpurun_out/kernels.log | cut -c1-600
timeout 600 $PT tests/test_api_gpu.py > gpurun_out/api.log 2>&1; echo "api rc=$?"; tail -n 3 gpurun_out/api.log | cut -c1-600
timeout 900 python scripts/profile_configs.py cfg2,cfg3,cfg5,cfg5_shard8 2>&1 | cut -c1-500
