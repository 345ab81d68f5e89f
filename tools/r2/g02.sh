#!/bin/bash
# round 2, call 2: new quadrant-block cost volume (parity + timing), full-size parity tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "split" > gpurun_out/r2_pytest_split.log 2>&1; tail -5 gpurun_out/r2_pytest_split.log
for v in quad scatter; do for f in dirty clean; do
  PWC_FLUSH=$f PWC_CV_SPLIT=$v timeout 60 python tools/cv_bench.py 8 20 split 2>&1 | tail -1
  PWC_FLUSH=$f PWC_CV_SPLIT=$v timeout 60 python tools/cv_bench.py 8 20 splitslot 2>&1 | tail -1
done; done
PWC_FLUSH=clean PWC_CV_SPLIT=quad timeout 60 python tools/cv_bench.py 32 10 split 2>&1 | tail -1
PWC_FLUSH=clean timeout 60 python tools/cv_bench.py 8 20 2>&1 | tail -1
PWC_FLUSH=clean timeout 60 python tools/cv_bench.py 8 20 slot 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s > gpurun_out/r2_pytest_fullsize.log 2>&1; grep -E "parity|passed|failed|Error|error" gpurun_out/r2_pytest_fullsize.log | head -30
