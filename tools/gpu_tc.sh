#!/bin/bash
mkdir -p gpurun_out
./tools/mma_bench.bin > gpurun_out/mma_bench.log 2>&1
timeout 300 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1
echo "rc=$?" >> gpurun_out/tc_probe.log
cat gpurun_out/mma_bench.log; cat gpurun_out/tc_probe.log | tail -40
