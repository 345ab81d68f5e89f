#!/bin/bash
timeout 300 python tools/halo_probe.py time 2>&1 | grep halo
PWC_HALO_STAGES=2 timeout 300 python tools/halo_probe.py time 2>&1 | grep halo | head -4
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "halo or f16 or head" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['gpu_launches'], d['steps'])"
