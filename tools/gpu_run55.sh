#!/bin/bash
CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/wtc_dbg.py 2>&1 | tail -8 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "tensor_core" 2>&1 | tail -15 | cut -c1-250
