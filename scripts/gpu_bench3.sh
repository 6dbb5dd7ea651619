#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | cut -c1-1800 | tee -a gpurun_out/summary.txt
}
short() { python - "$1" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print("value=%.0f ms/step=%.2f e2e=%.0f cov_ms=%.3f cov_frac=%.3f kernel_ms=%s clocks=%s" % (d['value'], d['ms_per_step'], d['e2e']['value'] or 0, r['avg_launch_ms'], r['frac'], {k:round(v,2) for k,v in r['kernel_ms_per_step'].items()}, d['clocks']))
PY
}
PT="python -m pytest -q --timeout 240 -p no:cacheprovider --tb=line"
run kernels 600 $PT tests/test_kernels_gpu.py
run api 900 $PT tests/test_api_gpu.py
for cfg in "4 0" "3 0" "2 0" "2 6" "3 4"; do
  set -- $cfg
  export OIVA_COV_STAGES=$1
  if [ "$2" != "0" ]; then export OIVA_COV_TEAMS=$2; else unset OIVA_COV_TEAMS; fi
  n=bench_s$1_t$2
  timeout 600 python bench.py --batch 256 --steps 3 --no-cpu > gpurun_out/$n.log 2>&1
  echo "=== $n rc=$?" | tee -a gpurun_out/summary.txt
  short gpurun_out/$n.log | tee -a gpurun_out/summary.txt
done
unset OIVA_COV_STAGES OIVA_COV_TEAMS
OIVA_SOLVER_ROWOWNER=1 timeout 600 python bench.py --batch 256 --steps 3 --no-cpu > gpurun_out/bench_rowowner.log 2>&1
echo "=== rowowner solver" | tee -a gpurun_out/summary.txt; short gpurun_out/bench_rowowner.log | tee -a gpurun_out/summary.txt
