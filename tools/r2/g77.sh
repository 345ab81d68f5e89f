cd /root/repo
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "programmatic" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
