cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "stride2_space" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
PWC_S2D=0 timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
