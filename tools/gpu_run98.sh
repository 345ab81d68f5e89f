#!/bin/bash
for s in 2 3 4; do echo "== PWC_HALO_STAGES=$s"; PWC_HALO_STAGES=$s timeout 200 python tools/halo_probe.py time 2>&1 | grep " halo" | grep -v " d[2-9]\| d16" | grep "64->32\|32->32\|16->16" ; done
timeout 200 python tools/halo_probe.py 2>&1 | tail -13
