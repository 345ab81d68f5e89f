cd /root/repo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:halo -c 40 --csv --log-file gpurun_out/r2_t4.csv python tools/s2d_once.py 16 32 16 192 512 > gpurun_out/r2_t4.log 2>&1; tail -2 gpurun_out/r2_t4.log; grep -v '^"' gpurun_out/r2_t4.csv | tail -3; grep '^"' gpurun_out/r2_t4.csv | cut -d, -f9,15 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:halo -c 40 --csv --log-file gpurun_out/r2_t5.csv python tools/s2d_once.py 16 32 16 224 512 > gpurun_out/r2_t5.log 2>&1; tail -2 gpurun_out/r2_t5.log; grep -v '^"' gpurun_out/r2_t5.csv | tail -3
timeout 600 compute-sanitizer --tool memcheck python tools/s2d_once.py 16 32 16 192 512 2>&1 | grep -v "cuKernelGetFunction\|Host Frame\|Saved host\|=========$" | head -30
