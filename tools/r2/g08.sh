#!/bin/bash
mkdir -p gpurun_out
for pitch in 88 96 152 160; do for e in 0 4; do
PWC_CV_EXP=$e PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 20 splitslot$pitch 2>&1 | tail -1 | sed "s/^/exp=$e /"
done; done
PWC_CV_EXP=4 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 splitslot152 2>&1 | tail -1
