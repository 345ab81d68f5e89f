#!/bin/bash
for w in 96 64 32 16; do echo "== PWC_HALO_MINW=$w"; PWC_HALO_MINW=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c40-130; done
