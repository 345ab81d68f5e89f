#!/bin/bash
mkdir -p gpurun_out
PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 3 slot > gpurun_out/cv_dbg31.log 2>&1; cat gpurun_out/cv_dbg31.log
