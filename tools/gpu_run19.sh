#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "cost_volume" > gpurun_out/pytest19.log 2>&1; tail -5 gpurun_out/pytest19.log
timeout 120 python tools/cv_bench.py 8 20 2>&1 | tail -2
PWC_CV_LEGACY=1 timeout 120 python tools/cv_bench.py 8 20 2>&1 | tail -1
timeout 120 python tools/cv_bench.py 32 10 2>&1 | tail -1
