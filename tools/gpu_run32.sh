#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "cost_volume" > gpurun_out/pytest32.log 2>&1; tail -5 gpurun_out/pytest32.log | cut -c1-250
PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 slot 2>&1 | head -22
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench32.log 2>&1
timeout 120 python tools/cv_bench.py 32 20 slot >> gpurun_out/cv_bench32.log 2>&1
timeout 120 python tools/cv_bench.py 8 20 slot >> gpurun_out/cv_bench32.log 2>&1
cat gpurun_out/cv_bench32.log
