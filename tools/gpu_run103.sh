#!/bin/bash
mkdir -p gpurun_out
timeout 40 ./tools/mma_sync_bench.bin > gpurun_out/mma_sync_bench.log 2>&1; cat gpurun_out/mma_sync_bench.log
