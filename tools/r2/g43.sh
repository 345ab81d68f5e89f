cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/halo_narrow_bench.py 2>&1 | tail -12
timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; tail -3 gpurun_out/r2_bench_f.err; cut -c1-200 gpurun_out/r2_bench_f.json
PWC_HALO_NO_NSPLIT=1 timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
