cd /root/repo
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "cost_volume" 2>&1 | tail -3
timeout 120 python tools/roofline_once.py 8 2>&1 | tail -1
PWC_CV_NO_TMA_OUT=1 timeout 120 python tools/roofline_once.py 8 2>&1 | tail -1
timeout 120 python tools/roofline_once.py 16 2>&1 | tail -1
timeout 120 python tools/roofline_once.py 32 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma-out', d['value'], d['burst_value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['probe']['sha256_16'])"
