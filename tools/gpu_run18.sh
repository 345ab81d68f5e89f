#!/bin/bash
for sh in 0 64 128 192 320 512; do echo "== shift $sh"; PWC_TC_SHIFT=$sh timeout 100 python tools/f16_probe.py 2>&1 | grep -E "^B2 14x32 Cin128\(cs128\) Cout128 d1 s1 scale1.0|^B1 28x64 Cin147" | cut -c1-100; done
