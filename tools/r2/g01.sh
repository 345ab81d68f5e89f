#!/bin/bash
# round 2, call 1: TMEM/FFMA2/store microbench, the never-executed row32 kernels (parity + timing)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_g01_smi.log 2>&1
timeout 120 ./tools/tmem_ld_bench.bin > gpurun_out/r2_tmem_ld_bench.log 2>&1; tail -45 gpurun_out/r2_tmem_ld_bench.log
PWC_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "row32" > gpurun_out/r2_pytest_row32.log 2>&1; tail -15 gpurun_out/r2_pytest_row32.log
for v in default row32 row32p; do
  PWC_CV_SPLIT=$v timeout 60 python tools/cv_bench.py 8 10 splitslot 2>&1 | tail -1
  PWC_CV_SPLIT=$v timeout 60 python tools/cv_bench.py 32 10 splitslot 2>&1 | tail -1
done
timeout 60 python tools/cv_bench.py 8 10 slot 2>&1 | tail -1
