#!/bin/bash
timeout 60 python tools/cv_bench.py 8 5 splitslot > gpurun_out/cv49_first.log 2>&1; echo "first rc=$?"; tail -2 gpurun_out/cv49_first.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "split" 2>&1 | tail -8 | cut -c1-250
timeout 120 python tools/cv_bench.py 8 20 split
timeout 120 python tools/cv_bench.py 8 20 splitslot
timeout 120 python tools/cv_bench.py 32 20 splitslot
PWC_CV_DEBUG=1 timeout 120 python tools/cv_bench.py 8 1 splitslot 2>&1 | sed -n 11,20p
