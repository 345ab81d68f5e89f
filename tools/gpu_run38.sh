#!/bin/bash
timeout 200 python tools/halo_probe.py 2>&1 | tail -16
timeout 300 python tools/halo_probe.py time 2>&1 | tail -8
