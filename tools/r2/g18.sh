#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/roofline_once.py 8 2>&1 | tail -3
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot152 2>&1 | tail -1
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2_pytest_all.log 2>&1; tail -30 gpurun_out/r2_pytest_all.log
