cd /root/repo
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_first -s 2 -c 1 -o gpurun_out/r2_ncu_conv_first -f python tools/fwd_once.py > gpurun_out/r2_ncu1.log 2>&1; tail -2 gpurun_out/r2_ncu1.log
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_halo -s 98 -c 1 -o gpurun_out/r2_ncu_halo_16 -f python tools/fwd_once.py > gpurun_out/r2_ncu2.log 2>&1; tail -2 gpurun_out/r2_ncu2.log
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_halo -s 135 -c 1 -o gpurun_out/r2_ncu_halo_128 -f python tools/fwd_once.py > gpurun_out/r2_ncu3.log 2>&1; tail -2 gpurun_out/r2_ncu3.log
ls -la gpurun_out/*.ncu-rep
