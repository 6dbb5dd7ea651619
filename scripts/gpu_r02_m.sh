#!/bin/bash
# Round-2 call M: determined (AuxIVA) batches M = K = 4 / 6 / 8, 256 mixtures: per-kernel times and ncu captures.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  echo "rc=$?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-3000 | tee -a gpurun_out/summary.txt
  tail -n 5 gpurun_out/$name.err | cut -c1-1500 | tee -a gpurun_out/summary.txt
}
run r02m_kernels 600 python scripts/profile_configs.py det4_b256,det6_b256,det8_b256
run r02m_ncu6 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_demix|k_ip_update" -s 8 -c 4 -o gpurun_out/r02m_det6 python scripts/profile_configs.py det6_b256
run r02m_ncu8 900 ncu --set full --clock-control none --import-source on -k regex:"k_cov|k_demix|k_ip_update" -s 10 -c 5 -o gpurun_out/r02m_det8 python scripts/profile_configs.py det8_b256
for n in det6 det8; do ncu -i gpurun_out/r02m_$n.ncu-rep --page raw --csv > gpurun_out/r02m_${n}_raw.csv 2>/dev/null; done
