#!/bin/bash
mkdir -p gpurun_out
P=29617
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench76_4gpu.json 2> gpurun_out/bench76_4gpu.err; cut -c1-200 gpurun_out/bench76_4gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --mode train --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench76_train_4gpu.json 2> gpurun_out/bench76_train_4gpu.err; cut -c1-230 gpurun_out/bench76_train_4gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --impl reference --gpus 4 --steps 2 --warmup 3 2>/dev/null | cut -c1-160
