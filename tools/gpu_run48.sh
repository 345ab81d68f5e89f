#!/bin/bash
PWC_CV_EXP=2 PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 splitslot 2>&1 | sed -n 11,20p
PWC_CV_EXP=6 PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 splitslot 2>&1 | sed -n 11,20p
