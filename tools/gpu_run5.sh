#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 2 -c 1 -o gpurun_out/tc_prof_3x -f python tools/tc_one.py 3 128 128 > gpurun_out/ncu_tc3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 2 -c 1 -o gpurun_out/tc_prof_1x -f python tools/tc_one.py 1 128 128 > gpurun_out/ncu_tc1.log 2>&1
tail -2 gpurun_out/ncu_tc3.log gpurun_out/ncu_tc1.log
