cd /root/repo
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err; tail -3 gpurun_out/r2_bench_h.err; cat gpurun_out/r2_bench_h.json
