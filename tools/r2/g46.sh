cd /root/repo
PWC_HALO_CORR2=1 timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
for args in "288 128 8 7 16" "16 16 16 224 512"; do
  PWC_HALO_CORR2=1 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14 | head -7
  PWC_HALO_CORR2=1 timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -2
done
PWC_HALO_CORR2=1 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_CORR2=1 timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
