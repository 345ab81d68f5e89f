cd /root/repo
for v in "PWC_S2D=0" "PWC_HALO_NO_TMA_STORE=1" "PWC_HALO_NO_NSPLIT=1" "PWC_NO_GRAPH=1"; do
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_t7.csv python tools/train_once.py 1 1 > gpurun_out/r2_t7.log 2>&1
  echo "$v: rows $(grep -c '^"' gpurun_out/r2_t7.csv) $(grep -v '^"' gpurun_out/r2_t7.csv | grep ERROR | head -1)"
done
