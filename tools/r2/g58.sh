cd /root/repo
PWC_SPLIT_ACT=0 PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2_t1.csv python tools/fwd_once.py > gpurun_out/r2_t1.log 2>&1; grep -v '^"' gpurun_out/r2_t1.csv | tail -3; grep -c halo gpurun_out/r2_t1.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:halo -c 40 --csv --log-file gpurun_out/r2_t2.csv python tools/halo_narrow_bench.py > gpurun_out/r2_t2.log 2>&1; grep -v '^"' gpurun_out/r2_t2.csv | tail -3; grep -c halo gpurun_out/r2_t2.csv
timeout 300 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_t3.csv python tools/train_once.py 8 1 > gpurun_out/r2_t3.log 2>&1; grep -v '^"' gpurun_out/r2_t3.csv | tail -3; grep -c halo gpurun_out/r2_t3.csv
timeout 600 compute-sanitizer --tool memcheck python tools/halo_narrow_dbg.py 32 32 16 96 256 2>&1 | tail -15
