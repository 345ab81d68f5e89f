cd /root/repo
PWC_WGRAD_TC_SMALL=1 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tc_small', d['value'], d['train']['value'], d['train']['ms_per_step'])"
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', d['value'], d['train']['value'], d['train']['ms_per_step'])"
