cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nsplit', d['value'], d['burst_value'], d['train']['value'], d['train']['ms_per_step'], d['probe']['sha256_16'])"
PWC_TC_NO_NSPLIT=1 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no nsplit', d['value'], d['burst_value'], d['train']['value'], d['train']['ms_per_step'], d['probe']['sha256_16'])"
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_f16_kernel -c 30 --csv --log-file gpurun_out/r2_f16.csv python tools/fwd_once.py > /dev/null 2>&1; grep tc_f16 gpurun_out/r2_f16.csv | cut -d, -f9,15 | tail -5
