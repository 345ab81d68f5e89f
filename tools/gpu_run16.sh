#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench16_2gpu.json 2> gpurun_out/bench16_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench16_ref.json 2> gpurun_out/bench16_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-pairs 5 > gpurun_out/bench16.json 2> gpurun_out/bench16.err
tail -2 gpurun_out/bench16_2gpu.err; python -c "
import json
for f in ['bench16_2gpu','bench16']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None, d.get('cpu_baseline'))
print(open('gpurun_out/bench16_ref.json').read()[:600])"
