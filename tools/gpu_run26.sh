#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench26_train.json 2> gpurun_out/bench26_train.err; cut -c1-600 gpurun_out/bench26_train.json; tail -3 gpurun_out/bench26_train.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches26_train.csv python tools/train_once.py 8 2 > gpurun_out/t26.log 2>&1; tail -2 gpurun_out/t26.log
