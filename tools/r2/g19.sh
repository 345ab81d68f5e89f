#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/roofline_once.py 8 2>&1 | tail -2
timeout 120 python tools/roofline_once.py 32 2>&1 | tail -1
timeout 300 python tools/dc_grad_dbg.py 2>&1 | tail -4
timeout 900 python -m pytest tests -q -m gpu --timeout 900 -k "split or shape_change or batched_weight or cli or guard or uint8" > gpurun_out/r2_pytest_sel.log 2>&1; tail -5 gpurun_out/r2_pytest_sel.log
