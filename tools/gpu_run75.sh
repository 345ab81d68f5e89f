#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>/dev/null | cut -c1-230
