cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
PWC_HALO_ROW64=1 timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
PWC_HALO_ROW64=1 timeout 300 python tools/halo_narrow_bench.py 2>&1 | head -2
timeout 300 python tools/halo_narrow_bench.py 2>&1 | head -2
PWC_HALO_ROW64=1 timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
PWC_HALO_ROW64=1 timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
