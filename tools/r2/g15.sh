#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "split" > gpurun_out/r2_pytest_split.log 2>&1; tail -2 gpurun_out/r2_pytest_split.log
PWC_WIDE=1 PWC_CV_DEBUG=1 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 3 splitslot152 2>&1 | grep -A9 "cv_quad dbg" | tail -9
for cfg in "1 152" "1 88" "0 148"; do set -- $cfg
PWC_WIDE=$1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot$2 2>&1 | tail -1 | sed "s/^/wide=$1 /"
done
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_EXP=2 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot152 2>&1 | tail -1
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 splitslot152 2>&1 | tail -1
