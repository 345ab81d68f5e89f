#!/bin/bash
for g in 1 2 4; do PWC_HALO_GROUP=$g timeout 120 python tools/halo_narrow_bench.py 2>&1 | tail -4; done
