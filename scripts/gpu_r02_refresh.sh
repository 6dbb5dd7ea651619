#!/bin/bash
# Short refresh of the bench evidence after a late change: tests, bench line, reference arm, launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/fin_pytest.log
timeout 1500 python bench.py > gpurun_out/fin_bench_n1.log 2> gpurun_out/fin_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/fin_refarm.log 2> gpurun_out/fin_refarm.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5 > gpurun_out/fin_ncu_list.log 2>&1; echo "ncu rc=$?"
cut -c1-400 gpurun_out/fin_bench_n1.log
timeout 600 python scripts/bench_configs.py --configs cfg1,cfg2,cfg3,cfg5 --cpu > gpurun_out/fin_configs.log 2> gpurun_out/fin_configs.err; echo "configs rc=$?"
timeout 600 python scripts/profile_configs.py cfg1,cfg2,cfg3,cfg5,cfg5_shard8,cfg4_b64,det4_b256,det6_b256,det8_b256 > gpurun_out/fin_kernels.log 2> gpurun_out/fin_kernels.err; echo "kernels rc=$?"
timeout 200 python scripts/trace_resident.py > gpurun_out/res_trace.jsonl 2> gpurun_out/res_trace.err; echo "trace rc=$?"
timeout 200 python scripts/ab_res_handover.py > gpurun_out/ab_res_handover.jsonl 2> gpurun_out/ab_res_handover.err; echo "ab rc=$?"
