#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "halo" 2>&1 | tail -6 | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c40-130
PWC_HALO_NO_FLAT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c40-130
