#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_quad -s 10 -c 1 -o gpurun_out/r2_cv_quad_final python tools/roofline_once.py 8 > gpurun_out/r2_ncu_cv_final.log 2>&1; tail -2 gpurun_out/r2_ncu_cv_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cost_volume_quad -c 60 --csv --log-file gpurun_out/r2_launches_cv_roofline.csv python tools/roofline_once.py 8 > /dev/null 2>&1; tail -3 gpurun_out/r2_launches_cv_roofline.csv | cut -c1-200
