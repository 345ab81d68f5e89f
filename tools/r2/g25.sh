#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2_pytest_all.log 2>&1; tail -12 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err; cut -c1-200 gpurun_out/r2_bench_d.json
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd2.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; tail -2 gpurun_out/r2_fwd_once.log
