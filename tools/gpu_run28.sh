#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -s > gpurun_out/pytest28.log 2>&1; tail -15 gpurun_out/pytest28.log | cut -c1-300
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench28_train.json 2> gpurun_out/bench28_train.err; cut -c1-300 gpurun_out/bench28_train.json; tail -3 gpurun_out/bench28_train.err
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches28_train.csv python tools/train_once.py 8 2 > gpurun_out/t28.log 2>&1; tail -2 gpurun_out/t28.log
