#!/bin/bash
mkdir -p gpurun_out

PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches27_train.csv python tools/train_once.py 8 2 > gpurun_out/t27.log 2>&1; tail -2 gpurun_out/t27.log
