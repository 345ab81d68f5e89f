#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/halo_probe.py time 2>&1 | tail -16
