#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/roofline_once.py 8 2>&1 | tail -2
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_pytest_all.log 2>&1; tail -6 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -3 gpurun_out/r2_bench_b.err; cut -c1-200 gpurun_out/r2_bench_b.json
