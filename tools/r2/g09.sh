#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "split" > gpurun_out/r2_pytest_split.log 2>&1; tail -2 gpurun_out/r2_pytest_split.log
PWC_CV_EXP=2 PWC_CV_DEBUG=1 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 3 splitslot152 2>&1 | grep -A9 "cv_quad dbg" | tail -9
for e in 0 4 2; do for pitch in 148 152; do
PWC_ROTATE=4 PWC_CV_EXP=$e PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot$pitch 2>&1 | tail -1 | sed "s/^/exp=$e /"
PWC_CV_EXP=$e PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot$pitch 2>&1 | tail -1 | sed "s/^/exp=$e /"
done; done
PWC_ROTATE=4 timeout 60 python tools/cv_bench.py 8 25 slot 2>&1 | tail -1
PWC_ROTATE=4 PWC_CV_SPLIT=scatter timeout 60 python tools/cv_bench.py 8 25 splitslot148 2>&1 | tail -1
