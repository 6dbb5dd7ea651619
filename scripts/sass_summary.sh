#!/bin/bash
# Per-object SASS mnemonic counts of the built library (cuobjdump -sass on the sm_100a objects): the evidence that the
# kernels use bulk TMA (UBLKCP), mbarriers (SYNCS), cluster multicast, fp64 FMA (DFMA/DMUL) and -- only in the
# microbenchmark -- the fp64 tensor-core instruction (DMMA).   usage: scripts/sass_summary.sh > profiles/sass_summary.txt
cd "$(dirname "$0")/../overiva_b200/csrc/build" || exit 1
printf "%-18s %8s %8s %8s %8s %8s %8s %8s %8s %9s\n" object UBLKCP SYNCS DFMA DMUL DMMA LDS.128 LDG REDG UCGABAR
for o in cov_m6.o cov_m8.o cov_m16.o stream_m6.o stream_m16.o resident_m4.o resident_m6.o resident_m8.o solve_tpb.o solve.o stft.o microbench.o; do
  [ -f $o ] || continue
  s=$(cuobjdump -sass $o)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-18s %8d %8d %8d %8d %8d %8d %8d %8d %9d\n" $o $(c UBLKCP) $(c SYNCS) $(c "DFMA") $(c "DMUL") $(c DMMA) $(c "LDS.128") $(c "LDG") $(c "REDG") $(c "UCGABAR")
done
echo
echo "multicast bulk copies (k_cov_tiled, cov_m16.o):"
cuobjdump -sass cov_m16.o | grep -E "UBLKCP" | sed 's/^[ \t]*//' | awk '{print $2, $3}' | sort | uniq -c
