#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/store_bw_bench.bin 2>&1 | tee gpurun_out/r2_store_bw_bench.log
