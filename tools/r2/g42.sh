cd /root/repo
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd3.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; tail -2 gpurun_out/r2_fwd_once.log
