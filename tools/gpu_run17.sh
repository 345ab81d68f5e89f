#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/f16_probe.py 2>&1 | grep -v "tf32x3" > gpurun_out/f16_probe17.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest17.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
PWC_TC_DEBUG=1 timeout 60 python tools/f16_dbg.py 2>&1 | grep -A17 "per-stage" | tail -19 > gpurun_out/f16_dbg17.log
cat gpurun_out/f16_probe17.log gpurun_out/pytest17.log; python -c "
import json
d=json.load(open('gpurun_out/bench17.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync_value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench17.err; cat gpurun_out/f16_dbg17.log
