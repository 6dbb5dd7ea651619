#!/bin/bash
# Short refresh of the bench evidence after a late change: tests, bench line, reference arm, launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/fin_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/fin_pytest.log
timeout 1500 python bench.py > gpurun_out/fin_bench_n1.log 2> gpurun_out/fin_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/fin_refarm.log 2> gpurun_out/fin_refarm.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 300 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --no-cfg5 > gpurun_out/fin_ncu_list.log 2>&1; echo "ncu rc=$?"
cut -c1-400 gpurun_out/fin_bench_n1.log
