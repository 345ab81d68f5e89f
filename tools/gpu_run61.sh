#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_halo -s 2 -c 1 -o gpurun_out/halo_prof -f python tools/halo_one.py > gpurun_out/ncu_halo.log 2>&1; tail -1 gpurun_out/ncu_halo.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 2 -c 1 -o gpurun_out/wgrad_tc_prof -f python tools/halo_one.py > gpurun_out/ncu_wtc.log 2>&1; tail -1 gpurun_out/ncu_wtc.log
