cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('side2', d['value'], d['burst_value'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'], d['probe']['sha256_16'])"
PWC_SIDE_SPLIT=0 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('inline', d['value'], d['burst_value'], d['e2e']['value'], d['train']['value'], d['train']['ms_per_step'], d['probe']['sha256_16'])"
