#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "cost_volume" > gpurun_out/pytest104.log 2>&1; tail -2 gpurun_out/pytest104.log
