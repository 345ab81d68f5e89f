#!/bin/bash
PWC_HALO_EXP=1 timeout 300 python tools/halo_probe.py time 2>&1 | grep "halo" | head -8
