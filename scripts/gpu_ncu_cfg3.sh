mkdir -p gpurun_out
OIVA_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none -k regex:"^k_cov$|^k_demix_power$" -s 6 -c 2 -o gpurun_out/cfg3 python scripts/profile_configs.py cfg3 > gpurun_out/cfg3_ncu.log 2>&1
ncu -i gpurun_out/cfg3.ncu-rep --page raw --csv > gpurun_out/cfg3_raw.csv 2>/dev/null; rm -f gpurun_out/cfg3.ncu-rep
tail -3 gpurun_out/cfg3_ncu.log | cut -c1-300
