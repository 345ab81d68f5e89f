cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv_first" 2>&1 | tail -3
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_first -c 3 --csv --log-file gpurun_out/r2_cf.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; grep conv_first gpurun_out/r2_cf.csv | cut -d, -f5,15-
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
