#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "cost_volume" > gpurun_out/pytest21.log 2>&1; tail -3 gpurun_out/pytest21.log
timeout 120 python tools/cv_bench.py 8 20 slot 2>&1 | tail -1
timeout 120 python tools/cv_bench.py 8 20 2>&1 | tail -1
timeout 120 python tools/cv_bench.py 32 10 slot 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_tma -c 1 -o gpurun_out/cv_tma2 -f python tools/cv_bench.py 8 2 slot > gpurun_out/ncu_cv_tma2.log 2>&1
