#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2; do
PWC_CV_EXP=$e PWC_CV_DEBUG=1 PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 3 splitslot 2>&1 | grep -A9 "cv_quad dbg" | tail -9
PWC_CV_EXP=$e PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 8 20 splitslot 2>&1 | tail -1
done
PWC_CV_SPLIT=quad timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --sampling-interval 0 --clock-control none --import-source on -k regex:cost_volume_quad -s 3 -c 6 -o gpurun_out/r2_cv_quad2 python tools/cv_bench.py 8 8 splitslot > gpurun_out/r2_ncu_cv_quad2.log 2>&1; tail -2 gpurun_out/r2_ncu_cv_quad2.log
