#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest39.log 2>&1; tail -3 gpurun_out/pytest39.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench39.json 2> gpurun_out/bench39.err; cut -c1-250 gpurun_out/bench39.json; tail -2 gpurun_out/bench39.err
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches39.csv python tools/fwd_once.py > gpurun_out/f39.log 2>&1; tail -2 gpurun_out/f39.log
