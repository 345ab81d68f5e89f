cd /root/repo
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_halo -s 98 -c 1 -o gpurun_out/r2_ncu_halo_16_final -f python tools/fwd_once.py > gpurun_out/r2_ncu2f.log 2>&1; tail -1 gpurun_out/r2_ncu2f.log
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_halo -s 135 -c 1 -o gpurun_out/r2_ncu_halo_128_final -f python tools/fwd_once.py > gpurun_out/r2_ncu3f.log 2>&1; tail -1 gpurun_out/r2_ncu3f.log
ls -la gpurun_out/*final.ncu-rep
