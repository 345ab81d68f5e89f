cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
for args in "288 128 8 7 16" "128 128 8 14 32"; do
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -14 | head -7
  timeout 60 python tools/halo_narrow_dbg.py $args 2>&1 | tail -2
done
timeout 900 python bench.py --no-train --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err; tail -3 gpurun_out/r2_bench_g.err; cut -c1-200 gpurun_out/r2_bench_g.json
PWC_HALO_NO_NSPLIT=1 timeout 900 python bench.py --no-train --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
