#!/bin/bash
mkdir -p gpurun_out
PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 slot 2>&1 | head -11
timeout 120 python tools/cv_bench.py 8 20 slot
PWC_CV_NOBACKOFF=1 timeout 120 python tools/cv_bench.py 8 20 slot
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "cost_volume" 2>&1 | tail -2
