#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_cv2.log
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench_ws.log 2>&1
timeout 120 python tools/cv_bench.py 16 20 >> gpurun_out/cv_bench_ws.log 2>&1
PWC_CV_NO_WS=1 timeout 120 python tools/cv_bench.py 8 20 >> gpurun_out/cv_bench_ws.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_r4 -s 3 -c 1 -o gpurun_out/cv_prof_ws -f python tools/cv_bench.py 8 3 > gpurun_out/ncu_cv_ws.log 2>&1
cat gpurun_out/pytest_cv2.log gpurun_out/cv_bench_ws.log
