#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu > gpurun_out/pytest25.log 2>&1; tail -60 gpurun_out/pytest25.log
