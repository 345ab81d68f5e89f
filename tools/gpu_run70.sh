#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "halo or f16 or head" 2>&1 | tail -2
timeout 300 python tools/halo_probe.py time 2>&1 | grep "halo" | head -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
