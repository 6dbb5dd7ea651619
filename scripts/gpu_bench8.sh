#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.log | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
short() { python - "$1" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print("value=%.0f ms/step=%.2f e2e=%.0f cov_ms=%.3f cov_frac=%.3f kernel_ms=%s launches=%d clocks=%s" % (d['value'], d['ms_per_step'], d['e2e']['value'] or 0, r['avg_launch_ms'], r['frac'], {k:round(v,2) for k,v in r['kernel_ms_per_step'].items()}, d['gpu_launches'], d['clocks']))
PY
}
PT="python -m pytest -q --timeout 240 -p no:cacheprovider --tb=short"
run kernels 600 $PT tests/test_kernels_gpu.py -k "fused or ip_update"
run api 900 $PT tests/test_api_gpu.py
for nf in 0 1 0; do
  OIVA_NO_FUSE=$nf timeout 900 python bench.py --no-cpu --steps 4 > gpurun_out/bench_nf$nf.log 2>&1; echo "=== bench no_fuse=$nf rc=$?" | tee -a gpurun_out/summary.txt; short gpurun_out/bench_nf$nf.log | tee -a gpurun_out/summary.txt
done
