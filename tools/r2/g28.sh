set -x
cd /root/repo
timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_EPI_DIRECT=1 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_EXP=2 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_SETS=2 timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
PWC_HALO_DEBUG=1 timeout 300 python tools/halo_narrow_dbg.py 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/fwd_once.py 2>&1 | tail -5
