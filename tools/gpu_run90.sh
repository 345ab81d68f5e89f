#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/pytest90.log 2>&1; tail -3 gpurun_out/pytest90.log
timeout 600 python bench.py --mode train --no-cpu-baseline > gpurun_out/bench90_train.json 2>gpurun_out/bench90.err; cut -c1-200 gpurun_out/bench90_train.json; tail -2 gpurun_out/bench90.err
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches90_train.csv python tools/train_once.py 8 2 > gpurun_out/t90.log 2>&1; tail -1 gpurun_out/t90.log
