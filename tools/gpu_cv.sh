#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "cost_volume or warp" 2>&1 | tail -5 > gpurun_out/pytest_cv.log
timeout 300 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench.log 2>&1
timeout 300 python tools/cv_bench.py 8 20 fused >> gpurun_out/cv_bench.log 2>&1
timeout 300 python tools/cv_bench.py 16 20 >> gpurun_out/cv_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_r4 -s 3 -c 1 -o gpurun_out/cv_prof2 -f python tools/cv_bench.py 8 3 > gpurun_out/ncu_cv2.log 2>&1
cat gpurun_out/pytest_cv.log gpurun_out/cv_bench.log
