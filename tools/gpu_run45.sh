#!/bin/bash
PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 splitslot 2>&1 | head -22
