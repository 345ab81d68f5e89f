#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_probe.py 2>&1 | head -24 > gpurun_out/tc_probe2.log
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "precision|passed|failed|FAILED|Error|error" | tail -40 > gpurun_out/pytest2.log
for prec in 3xtf32; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err
done
cat gpurun_out/tc_probe2.log gpurun_out/pytest2.log; for prec in 3xtf32; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$prec.json')); print('$prec', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/bench_$prec.err; done
