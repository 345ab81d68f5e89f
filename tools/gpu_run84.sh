#!/bin/bash
timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>/dev/null | cut -c70-160
PWC_WGRAD_TC_SMALL=1 timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>/dev/null | cut -c70-160
