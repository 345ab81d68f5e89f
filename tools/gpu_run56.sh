#!/bin/bash
timeout 300 compute-sanitizer --tool memcheck python tools/wtc_dbg.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         in" | head -40 | cut -c1-220
