#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench66.json 2> gpurun_out/bench66.err; cut -c1-200 gpurun_out/bench66.json
timeout 600 python bench.py --batch 32 --steps 10 --no-cpu-baseline > gpurun_out/bench66_b32.json 2>/dev/null; cut -c1-200 gpurun_out/bench66_b32.json
timeout 900 python bench.py --mode train > gpurun_out/bench66_train.json 2>/dev/null; cut -c1-230 gpurun_out/bench66_train.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench66_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench66_ref.json
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches66_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b66.log 2>&1; tail -1 gpurun_out/b66.log | cut -c1-100
