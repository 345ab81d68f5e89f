cd /root/repo
timeout 300 python tools/train_once.py 8 3 2>&1 | tail -3
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/train_once.py 8 2 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_train.csv python tools/train_once.py 8 3 > gpurun_out/r2_train_once.log 2>&1; tail -2 gpurun_out/r2_train_once.log; grep -v '^"' gpurun_out/r2_launches_train.csv | tail -4
