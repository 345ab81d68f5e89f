cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -60
