cd /root/repo
timeout 600 python bench.py --batch 32 --no-train --no-cpu-baseline > gpurun_out/r2_bench_b32.json 2> gpurun_out/r2_bench_b32.err; tail -2 gpurun_out/r2_bench_b32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b32.json').read().strip().splitlines()[-1])
print('B=32 value',d['value'],'ms',d['ms_per_step'],'burst',d['burst_value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['us_per_launch'],'conv',d['roofline_conv']['frac'],d['clocks'])
PY
timeout 600 python bench.py --batch 16 --no-train --no-cpu-baseline > gpurun_out/r2_bench_b16.json 2> gpurun_out/r2_bench_b16.err; tail -2 gpurun_out/r2_bench_b16.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b16.json').read().strip().splitlines()[-1])
print('B=16 value',d['value'],'ms',d['ms_per_step'],'burst',d['burst_value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['us_per_launch'],'conv',d['roofline_conv']['frac'],d['clocks'])
PY
