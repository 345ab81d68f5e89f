#!/bin/bash
mkdir -p gpurun_out
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:cost_volume_quad -s 6 -c 2 -o gpurun_out/r2_cv_quad3 python tools/cv_bench.py 8 2 splitslot152 > gpurun_out/r2_ncu_cv_quad3.log 2>&1; tail -2 gpurun_out/r2_ncu_cv_quad3.log
