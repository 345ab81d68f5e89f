cd /root/repo
PWC_HALO_EXP=2 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no-store', d['value'], d['train']['value'], d['train']['ms_per_step'])"
