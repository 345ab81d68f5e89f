#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -5 gpurun_out/r2_bench_2gpu.err; cut -c1-300 gpurun_out/r2_bench_2gpu.json
