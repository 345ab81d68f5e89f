#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/halo_probe.py time 2>&1 | grep -E "halo|stream" | head -16
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>/dev/null | cut -c1-230
timeout 120 python tools/cv_bench.py 8 20 splitslot
