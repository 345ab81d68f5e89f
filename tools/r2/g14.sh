#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "split" > gpurun_out/r2_pytest_split.log 2>&1; tail -2 gpurun_out/r2_pytest_split.log
for cfg in "88 152" "96 160" "96 96" "96 128"; do set -- $cfg
PWC_WIDE=$1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot$2 2>&1 | tail -1 | sed "s/^/head=$1 /"
done
PWC_WIDE=96 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 splitslot160 2>&1 | tail -1
