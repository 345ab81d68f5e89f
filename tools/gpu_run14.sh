#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 3" "2 1" "3 1" "2 3"; do set -- $cfg; echo "== stages $1 nmain $2"; PWC_TC_PREFETCH=0 PWC_TC_STAGES=$1 PWC_TC_NMAIN=$2 timeout 120 python tools/f16_probe.py 2>&1 | grep "^time" | grep " f16" ; done > gpurun_out/occ_probe.log 2>&1
cat gpurun_out/occ_probe.log
