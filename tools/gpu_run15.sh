#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "precision|passed|failed|FAILED|Error" | tail -12 > gpurun_out/pytest15.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 225 -c 75 --csv --log-file gpurun_out/launches15.csv python tools/fwd_once.py 8 4 > gpurun_out/ncu15.log 2>&1
cat gpurun_out/pytest15.log; python -c "
import json
d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync_value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench15.err
