#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/bench_configs.py --cpu > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
echo "rc=$?"; cut -c1-900 gpurun_out/configs.jsonl; tail -n 5 gpurun_out/configs.err
timeout 600 python scripts/bench_freq_sharded.py > gpurun_out/cfg5_n1.json 2> gpurun_out/cfg5_n1.err; echo "rc=$?"; cat gpurun_out/cfg5_n1.json; tail -n 3 gpurun_out/cfg5_n1.err
