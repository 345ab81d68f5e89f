#!/bin/bash
for e in 0 1 2 4 5 7; do echo "== PWC_CV_EXP=$e"; PWC_CV_EXP=$e timeout 120 python tools/cv_bench.py 8 20 splitslot 2>&1 | tail -1; done
PWC_CV_EXP=1 PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 splitslot 2>&1 | head -10
