#!/bin/bash
# first GPU pass: tests, smoke, bench, ncu of the cost volume
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 300 python tools/cv_bench.py 8 20 > gpurun_out/cv_bench.log 2>&1
timeout 300 python tools/cv_bench.py 8 20 fused >> gpurun_out/cv_bench.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-pairs 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_volume_r4 -s 3 -c 2 -o gpurun_out/cv_prof -f python tools/cv_bench.py 8 3 > gpurun_out/ncu_cv.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/cv_bench.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
