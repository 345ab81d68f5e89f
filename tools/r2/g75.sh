cd /root/repo
PWC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_first -s 2 -c 1 -o gpurun_out/r2_ncu_conv_first2 -f python tools/fwd_once.py > gpurun_out/r2_ncu1.log 2>&1; tail -1 gpurun_out/r2_ncu1.log
