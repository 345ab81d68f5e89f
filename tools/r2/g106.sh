cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --mode train --steps 20 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'])"
