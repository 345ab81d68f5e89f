#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/r2_pytest_all.log 2>&1; tail -25 gpurun_out/r2_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -3 gpurun_out/r2_bench_a.err; cut -c1-400 gpurun_out/r2_bench_a.json
