#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rgb16 -c 3 --csv --log-file gpurun_out/rgb16.csv python tools/fwd_once.py > /dev/null 2>&1; grep rgb16 gpurun_out/rgb16.csv | tail -1 | cut -c1-300
