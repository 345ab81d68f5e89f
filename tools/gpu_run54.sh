#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "tensor_core" 2>&1 | tail -15 | cut -c1-250
