cd /root/repo
PWC_WGRAD_STREAM=1 timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
PWC_WGRAD_STREAM=1 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ws', d['value'], d['train']['value'], d['train']['ms_per_step'])"
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('main', d['value'], d['train']['value'], d['train']['ms_per_step'])"
