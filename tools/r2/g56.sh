cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_train.csv python tools/train_once.py 8 3 > gpurun_out/r2_train_once.log 2>&1; tail -2 gpurun_out/r2_train_once.log
