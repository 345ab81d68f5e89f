cd /root/repo
timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; tail -3 gpurun_out/r2_bench_e.err; cut -c1-400 gpurun_out/r2_bench_e.json
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
