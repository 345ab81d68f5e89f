#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; tail -3 gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_4gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], d['train']['ms_per_step'], 'exposed', d['train'].get('allreduce_exposed_ms'), d['probe'], d['clocks'])
PY
