#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./tools/umma_bench.bin > gpurun_out/umma_bench.log 2>&1; cat gpurun_out/umma_bench.log
