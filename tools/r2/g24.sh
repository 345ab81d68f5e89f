#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; tail -3 gpurun_out/r2_bench_8gpu.err; cut -c1-250 gpurun_out/r2_bench_8gpu.json
