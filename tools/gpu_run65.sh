#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "flow_head" 2>&1 | tail -2
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches65.csv python tools/fwd_once.py > gpurun_out/f65.log 2>&1; tail -1 gpurun_out/f65.log
