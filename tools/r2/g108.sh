cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/ddp_grad_check.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu_final2.json 2> gpurun_out/r2_bench_2gpu_final2.err; tail -2 gpurun_out/r2_bench_2gpu_final2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_final2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], d['train']['ms_per_step'], 'exposed', d['train'].get('allreduce_exposed_ms'), d['probe']['sha256_16'], d['clocks'])
PY
