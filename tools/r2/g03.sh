#!/bin/bash
mkdir -p gpurun_out
PWC_CV_DEBUG=1 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 3 splitslot 2>&1 | tail -24
PWC_CV_SPLIT=quad timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_volume_quad -s 3 -c 1 -o gpurun_out/r2_cv_quad python tools/cv_bench.py 8 3 splitslot > gpurun_out/r2_ncu_cv_quad.log 2>&1; tail -3 gpurun_out/r2_ncu_cv_quad.log
