#!/bin/bash
mkdir -p gpurun_out
python -m pwcnet_b200.build > /dev/null 2>&1
for m in 0 1; do echo "== PWC_HALO_DESC=$m"; PWC_HALO_DESC=$m timeout 120 python tools/halo_probe.py 2>&1 | tail -9; done
