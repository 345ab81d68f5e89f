#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 213 -c 71 --csv --log-file gpurun_out/launches8.csv python tools/fwd_once.py 8 4 > gpurun_out/ncu8.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench8.err
