#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest10.log
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench10.log 2>&1
timeout 120 python tools/cv_bench.py 8 20 fused >> gpurun_out/cv_bench10.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err
cat gpurun_out/pytest10.log gpurun_out/cv_bench10.log; python -c "
import json
d=json.load(open('gpurun_out/bench10.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync_value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench10.err
