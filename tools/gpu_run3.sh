#!/bin/bash
mkdir -p gpurun_out
# 2 image copies + 69 kernels per forward; skip 2 forwards
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 142 -c 71 --csv --log-file gpurun_out/launches_fwd.csv python tools/fwd_once.py 8 3 > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/ncu_fwd.log
