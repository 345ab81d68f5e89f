cd /root/repo
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "cost_volume" 2>&1 | tail -3
timeout 120 python tools/roofline_once.py 8 2>&1 | tail -2
PWC_CV_NO_ST256=1 timeout 120 python tools/roofline_once.py 8 2>&1 | tail -1
timeout 120 python tools/roofline_once.py 16 2>&1 | tail -1
timeout 120 python tools/roofline_once.py 32 2>&1 | tail -1
