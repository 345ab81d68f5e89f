cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "tcgen05 or halo or conv" 2>&1 | tail -3
echo "--- split-K on"; PWC_TC_KSPLIT=1 timeout 120 python tools/ksplit_once.py 2>&1 | tail -4
echo "--- split-K off"; timeout 120 python tools/ksplit_once.py 2>&1 | tail -4
