#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest24.log 2>&1; tail -3 gpurun_out/pytest24.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench24.json 2> gpurun_out/bench24.err; cut -c1-300 gpurun_out/bench24.json
timeout 120 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench24.log 2>&1
timeout 120 python tools/cv_bench.py 32 20 >> gpurun_out/cv_bench24.log 2>&1
timeout 120 python tools/cv_bench.py 8 20 slot >> gpurun_out/cv_bench24.log 2>&1
cat gpurun_out/cv_bench24.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_tma -s 3 -c 1 -o gpurun_out/cv_tma_prof -f python tools/cv_bench.py 8 3 > gpurun_out/ncu_cv24.log 2>&1
tail -3 gpurun_out/ncu_cv24.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches24.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b24.log 2>&1
