#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_split -s 3 -c 1 -o gpurun_out/cv_split_prof -f python tools/cv_bench.py 8 3 splitslot > gpurun_out/ncu_cv46.log 2>&1
tail -2 gpurun_out/ncu_cv46.log
