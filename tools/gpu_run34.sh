#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest34.log 2>&1; tail -3 gpurun_out/pytest34.log
