#!/bin/bash
for s in 8 4 2 1; do echo "== PWC_HALO_SETS=$s"; PWC_HALO_SETS=$s timeout 200 python tools/halo_probe.py time 2>&1 | grep " halo" | grep -v " d[2-9]\| d16" ; done
