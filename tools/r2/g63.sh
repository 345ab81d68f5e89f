cd /root/repo
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_train.csv python tools/train_once.py 8 3 > gpurun_out/r2_train_once.log 2>&1; tail -1 gpurun_out/r2_train_once.log
