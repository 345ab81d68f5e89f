#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
PWC_TC_NO_DEEP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -2
