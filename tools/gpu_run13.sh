#!/bin/bash
mkdir -p gpurun_out
for pf in 0 4 8 16; do echo "== prefetch $pf"; PWC_TC_PREFETCH=$pf timeout 120 python tools/f16_probe.py 2>&1 | grep "^time" | grep -E "128->128 d1|147->128|32->32|16->16"; done > gpurun_out/pf_probe.log 2>&1
PWC_TC_DEBUG=1 timeout 60 python tools/f16_dbg.py 2>&1 | grep -A17 "per-stage" | tail -18 >> gpurun_out/pf_probe.log
cat gpurun_out/pf_probe.log
