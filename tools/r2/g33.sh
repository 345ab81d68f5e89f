cd /root/repo
PWC_HALO_EXP=1 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
for args in "16 16 16 224 512" "128 128 8 112 256"; do
  PWC_HALO_EXP=1 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -12
done
