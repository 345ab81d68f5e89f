#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_tc -s 3 -c 1 -o gpurun_out/cv_tc_prof -f python tools/cv_bench.py 8 3 slot > gpurun_out/ncu_cv30.log 2>&1
tail -3 gpurun_out/ncu_cv30.log
