#!/bin/bash
# Round-2 call F: tiled covariance variants (frames per stage x mid-chunk polling), all from one source tree.
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
TAILN=2 run r02f_tiled 240 python scripts/check_tiled.py
run r02f_default 300 python scripts/profile_configs.py cfg5,cfg5_shard8
for v in nomid tc4 tc4nomid; do
  export OVERIVA_B200_LIB=$PWD/overiva_b200/lib/variants/liboveriva_b200_$v.so
  TAILN=2 run r02f_${v}_check 240 python scripts/check_tiled.py
  run r02f_$v 300 python scripts/profile_configs.py cfg5,cfg5_shard8
  for S in 4 8; do OIVA_COV_TILED_STAGES=$S run r02f_${v}_S$S 300 python scripts/profile_configs.py cfg5; done
done
unset OVERIVA_B200_LIB
