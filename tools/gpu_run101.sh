#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/pytest101.log 2>&1; tail -5 gpurun_out/pytest101.log
timeout 600 python bench.py --mode train --no-cpu-baseline > gpurun_out/bench101_train.json 2>gpurun_out/bench101.err; python -c "
import json; d=json.load(open('gpurun_out/bench101_train.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('sync_value'), d['gpu_launches'])"; tail -2 gpurun_out/bench101.err
