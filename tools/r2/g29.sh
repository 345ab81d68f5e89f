cd /root/repo
for args in "16 16 16 224 512" "128 128 8 112 256"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -12
  PWC_HALO_EPI_DIRECT=1 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -12
  PWC_HALO_EXP=2 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -12
done
