cd /root/repo
for args in "16 16 16 224 512" "128 128 8 112 256" "96 64 8 112 256"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv
