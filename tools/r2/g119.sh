cd /root/repo
timeout 300 python -m pytest tests/test_gpu_train.py -x -q -k "uncleared" 2>&1 | tail -3
