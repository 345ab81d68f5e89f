#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_bwd_tiled -s 2 -c 1 -o gpurun_out/cb94 -f python tools/ws_probe.py > gpurun_out/ncu94.log 2>&1; tail -2 gpurun_out/ncu94.log
