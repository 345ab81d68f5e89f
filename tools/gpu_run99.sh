#!/bin/bash
for s in 1 2 4; do echo "== PWC_HALO_SETS=$s (stages auto)"; PWC_HALO_SETS=$s timeout 200 python tools/halo_probe.py time 2>&1 | grep " halo" | grep -v " d[2-9]\| d16" | grep "64->32\|32->32\|16->16\|96->64" ; done
