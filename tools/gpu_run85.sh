#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest85.log 2>&1; tail -2 gpurun_out/pytest85.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench85.json 2> gpurun_out/bench85.err; cut -c1-200 gpurun_out/bench85.json
timeout 600 python bench.py --batch 32 --steps 10 --no-cpu-baseline > gpurun_out/bench85_b32.json 2>/dev/null; cut -c40-130 gpurun_out/bench85_b32.json
timeout 900 python bench.py --mode train > gpurun_out/bench85_train.json 2>/dev/null; cut -c70-160 gpurun_out/bench85_train.json
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches85_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b85.log 2>&1
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches85_train.csv python tools/train_once.py 8 2 > gpurun_out/t85.log 2>&1; tail -1 gpurun_out/t85.log
