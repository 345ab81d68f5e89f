#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench87.json 2> gpurun_out/bench87.err; python -c "
import json; d=json.load(open('gpurun_out/bench87.json')); print(d['value'], d['roofline_conv'])"; tail -2 gpurun_out/bench87.err
