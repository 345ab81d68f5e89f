cd /root/repo
for args in "16 16 16 224 512" "32 32 16 112 256"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14
done
