#!/bin/bash
mkdir -p gpurun_out
for mt in 1 2; do
echo "== MT $mt" >> gpurun_out/tc_probe7.log
PWC_TC_MT=$mt timeout 200 python tools/tc_probe.py 2>&1 | grep -v "split0" | cut -c1-110 >> gpurun_out/tc_probe7.log
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest7.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
cat gpurun_out/tc_probe7.log gpurun_out/pytest7.log; python -c "
import json
d=json.load(open('gpurun_out/bench7.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; tail -3 gpurun_out/bench7.err
