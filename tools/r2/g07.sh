#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "split" > gpurun_out/r2_pytest_split.log 2>&1; tail -3 gpurun_out/r2_pytest_split.log
for e in 0 1 2; do
PWC_CV_EXP=$e PWC_CV_DEBUG=1 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 3 splitslot 2>&1 | grep -A9 "cv_quad dbg" | tail -9
PWC_CV_EXP=$e PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 20 splitslot 2>&1 | tail -1
done
PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 splitslot 2>&1 | tail -1
PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 20 splitslot84 2>&1 | tail -1
