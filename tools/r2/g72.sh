cd /root/repo
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err; cut -c1-260 gpurun_out/r2_bench_final.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; cut -c1-400 gpurun_out/r2_bench_final_ref.json
PWC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_fwd_final.csv python tools/fwd_once.py > gpurun_out/r2_fwd_once.log 2>&1; tail -1 gpurun_out/r2_fwd_once.log
PWC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_train_final.csv python tools/train_once.py 8 3 > gpurun_out/r2_train_once.log 2>&1; tail -1 gpurun_out/r2_train_once.log
