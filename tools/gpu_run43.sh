#!/bin/bash
mkdir -p gpurun_out
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench43_2gpu.json 2> gpurun_out/bench43_2gpu.err; cut -c1-300 gpurun_out/bench43_2gpu.json; tail -2 gpurun_out/bench43_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --mode train --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench43_train_2gpu.json 2> gpurun_out/bench43_train_2gpu.err; cut -c1-300 gpurun_out/bench43_train_2gpu.json; tail -2 gpurun_out/bench43_train_2gpu.err
