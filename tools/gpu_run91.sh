#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ws_probe.py 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_small -s 2 -c 1 -o gpurun_out/ws91 -f python tools/ws_probe.py > gpurun_out/ncu91.log 2>&1; tail -2 gpurun_out/ncu91.log
ls -la gpurun_out/ws91.ncu-rep
