#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest37.log 2>&1; tail -3 gpurun_out/pytest37.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench37.json 2> gpurun_out/bench37.err; cut -c1-250 gpurun_out/bench37.json; tail -2 gpurun_out/bench37.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench37_train.json 2> gpurun_out/bench37_train.err; cut -c1-250 gpurun_out/bench37_train.json
