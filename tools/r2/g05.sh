#!/bin/bash
mkdir -p gpurun_out
for pitch in 84 96 128 148 160; do
PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 20 splitslot$pitch 2>&1 | tail -1
done
PWC_CV_SPLIT=scatter timeout 60 python tools/cv_bench.py 8 20 splitslot84 2>&1 | tail -1
