cd /root/repo
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_final7.json 2> gpurun_out/r2_bench_final7.err; tail -2 gpurun_out/r2_bench_final7.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final7.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'burst',d['burst_value'],'e2e',d['e2e']['value'],'sync',d['e2e']['sync_value'])
print('roofline',d['roofline']['frac'],d['roofline']['us_per_launch'],'conv',d['roofline_conv']['frac'],d['roofline_conv']['mma_issue_frac'],d['roofline_conv']['us_per_launch'])
print('train',d['train']['value'],d['train']['ms_per_step'],'cpu',d['cpu_baseline']['value'],'clocks',d['clocks'],d['probe']['sha256_16'])
PY
