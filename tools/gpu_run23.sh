#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest23.log 2>&1; tail -3 gpurun_out/pytest23.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench23.json 2> gpurun_out/bench23.err; cat gpurun_out/bench23.json | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/bench23_b32.json 2> gpurun_out/bench23_b32.err; cat gpurun_out/bench23_b32.json | cut -c1-400
