cd /root/repo
for args in "288 128 8 7 16" "224 128 8 14 32" "128 128 8 14 32" "128 96 8 7 16"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14
  PWC_HALO_NO_NSPLIT=1 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -4
done
