#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest41.log 2>&1; tail -3 gpurun_out/pytest41.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench41.json 2> gpurun_out/bench41.err; cut -c1-250 gpurun_out/bench41.json; tail -2 gpurun_out/bench41.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_train.json 2> gpurun_out/bench41_train.err; cut -c1-250 gpurun_out/bench41_train.json
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches41_train.csv python tools/train_once.py 8 2 > gpurun_out/t41.log 2>&1; tail -1 gpurun_out/t41.log
