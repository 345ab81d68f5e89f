cd /root/repo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_t6.csv python tools/train_once.py 1 1 > gpurun_out/r2_t6.log 2>&1; tail -1 gpurun_out/r2_t6.log; grep -v '^"' gpurun_out/r2_t6.csv | tail -3; grep -c '^"' gpurun_out/r2_t6.csv
timeout 1200 compute-sanitizer --tool memcheck python tools/train_once.py 1 1 2>&1 | grep -v "cuKernelGetFunction\|Host Frame\|Saved host\|=========$" | head -40
