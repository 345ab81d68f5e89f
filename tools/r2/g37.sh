cd /root/repo
timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_NO_CORR2=1 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
for args in "16 16 16 224 512" "32 32 16 112 256"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14
done
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/fwd_once.py 2>&1 | tail -8
