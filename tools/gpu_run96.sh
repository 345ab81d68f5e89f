#!/bin/bash
PWC_HALO_DEBUG=1 timeout 100 python tools/halo_dbg.py 16 224 512 16 16 2>&1 | tail -9
PWC_HALO_DEBUG=1 timeout 100 python tools/halo_dbg.py 16 112 256 32 32 2>&1 | tail -9
