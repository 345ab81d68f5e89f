#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/cv_bench.py 8 5 split > gpurun_out/cv44_first.log 2>&1; echo "first rc=$?"; tail -2 gpurun_out/cv44_first.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "split" 2>&1 | tail -12 | cut -c1-200
timeout 120 python tools/cv_bench.py 8 20 split
timeout 120 python tools/cv_bench.py 8 20 splitslot
timeout 120 python tools/cv_bench.py 32 20 splitslot
