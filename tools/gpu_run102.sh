#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest102.log 2>&1; tail -2 gpurun_out/pytest102.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench102.json 2> gpurun_out/bench102.err; cut -c1-160 gpurun_out/bench102.json
timeout 900 python bench.py --mode train > gpurun_out/bench102_train.json 2>/dev/null; cut -c70-170 gpurun_out/bench102_train.json
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches102_train.csv python tools/train_once.py 8 2 > gpurun_out/t102.log 2>&1; tail -1 gpurun_out/t102.log
