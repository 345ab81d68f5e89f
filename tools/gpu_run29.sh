#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/cv_bench.py 8 5 > gpurun_out/cv29_first.log 2>&1; echo "first rc=$?"; tail -3 gpurun_out/cv29_first.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "cost_volume" > gpurun_out/pytest29.log 2>&1; tail -15 gpurun_out/pytest29.log | cut -c1-250
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench29.log 2>&1
timeout 120 python tools/cv_bench.py 32 20 >> gpurun_out/cv_bench29.log 2>&1
timeout 120 python tools/cv_bench.py 8 20 slot >> gpurun_out/cv_bench29.log 2>&1
PWC_CV_KERNEL=tma timeout 120 python tools/cv_bench.py 8 20 slot >> gpurun_out/cv_bench29.log 2>&1
cat gpurun_out/cv_bench29.log
