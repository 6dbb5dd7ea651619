#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q --timeout 300 -p no:cacheprovider --tb=short"
timeout 1200 $PT tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 900 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3 > gpurun_out/configs_graph.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/configs_graph.jsonl'):
    d=json.loads(l); print(d['config'], 'ms', round(d['ms_per_call_device_resident'],3), 'mix-s/s', round(d['mixture_s_per_s']), 'numpy io ms', round(d.get('ms_per_call_numpy_in_out',0),2))
PY
OIVA_NO_GRAPH=1 timeout 900 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3 > gpurun_out/configs_nograph.jsonl 2>> gpurun_out/configs.err; python - <<'PY'
import json
for l in open('gpurun_out/configs_nograph.jsonl'):
    d=json.loads(l); print('nograph', d['config'], 'ms', round(d['ms_per_call_device_resident'],3))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -s 264 -c 176 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
