#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/f16_probe.py > gpurun_out/f16_probe.log 2>&1
echo "rc=$?" >> gpurun_out/f16_probe.log
cat gpurun_out/f16_probe.log
