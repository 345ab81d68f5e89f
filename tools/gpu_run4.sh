#!/bin/bash
mkdir -p gpurun_out
for cs in 1 2 4; do
echo "== cluster $cs" >> gpurun_out/tc_probe4.log
PWC_TC_CLUSTER=$cs timeout 120 python tools/tc_probe.py 2>&1 | grep -v "split0" >> gpurun_out/tc_probe4.log
echo "rc=$?" >> gpurun_out/tc_probe4.log
done
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench4.log 2>&1
cat gpurun_out/tc_probe4.log gpurun_out/cv_bench4.log
