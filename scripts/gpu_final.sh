#!/bin/bash
# Round-end style record: full GPU test suite, default bench, ncu launch list of the same command, config table.
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 300 -p no:cacheprovider --tb=short"
timeout 1200 $PT tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 350 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_ip_update_tpb|k_demix_power" -s 190 -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 python scripts/bench_configs.py --cpu > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cut -c1-330 gpurun_out/configs.jsonl
