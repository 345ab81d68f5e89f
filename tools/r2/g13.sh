#!/bin/bash
mkdir -p gpurun_out
for pitch in 88 96 152; do
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 25 splitslot$pitch 2>&1 | tail -1
done
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 16 15 splitslot88 2>&1 | tail -1
PWC_WIDE=1 PWC_ROTATE=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 splitslot88 2>&1 | tail -1
