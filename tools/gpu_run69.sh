#!/bin/bash
timeout 600 python bench.py --mode train --steps 5 --no-cpu-baseline 2>&1 | tail -8 | cut -c1-300
