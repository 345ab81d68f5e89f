cd /root/repo
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe', d['value'], d['burst_value'], d['train']['value'], d['train']['ms_per_step'], d['probe']['sha256_16'])"
