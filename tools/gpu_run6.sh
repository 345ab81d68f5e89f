#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest6.log; timeout 200 python tools/tc_probe.py 2>&1 | grep "^time" | grep -v split0 > gpurun_out/tc_time6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/launches6.csv python tools/fwd_once.py 8 4 > gpurun_out/ncu6.log 2>&1
cat gpurun_out/pytest6.log gpurun_out/tc_time6.log; python -c "
import json
d=json.load(open('gpurun_out/bench6.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench6.err
