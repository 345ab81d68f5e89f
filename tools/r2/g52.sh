cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd3.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; tail -2 gpurun_out/r2_fwd_once.log
