#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_pytest_all.log 2>&1; tail -8 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -3 gpurun_out/r2_bench_c.err; cut -c1-200 gpurun_out/r2_bench_c.json
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; tail -2 gpurun_out/r2_fwd_once.log
