#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest64.log 2>&1; tail -3 gpurun_out/pytest64.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench64.json 2>gpurun_out/bench64.err; cut -c1-200 gpurun_out/bench64.json; tail -2 gpurun_out/bench64.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-230
