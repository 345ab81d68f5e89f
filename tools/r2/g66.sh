cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k training 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma-dgrad', d['value'], d['train']['value'], d['train']['ms_per_step'])"
PWC_HALO_NO_TMA_DGRAD=1 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('reg-dgrad', d['value'], d['train']['value'], d['train']['ms_per_step'])"
