#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/pytest95.log 2>&1; tail -3 gpurun_out/pytest95.log
timeout 300 python tools/ws_probe.py 2>&1 | tail -5
timeout 600 python bench.py --mode train --no-cpu-baseline > gpurun_out/bench95_train.json 2>gpurun_out/bench95.err; cut -c1-200 gpurun_out/bench95_train.json; tail -2 gpurun_out/bench95.err
