#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest100.log 2>&1; tail -2 gpurun_out/pytest100.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench100.json 2> gpurun_out/bench100.err; cut -c1-200 gpurun_out/bench100.json
timeout 900 python bench.py --mode train > gpurun_out/bench100_train.json 2>/dev/null; cut -c70-170 gpurun_out/bench100_train.json
