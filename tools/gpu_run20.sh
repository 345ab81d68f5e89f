#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_tma -c 2 -o gpurun_out/cv_tma1 -f python tools/cv_bench.py 8 3 > gpurun_out/ncu_cv_tma1.log 2>&1
tail -3 gpurun_out/ncu_cv_tma1.log
