cd /root/repo
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q -k "split_k_cluster" > gpurun_out/r2_sanitizer_ksplit.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_sanitizer_ksplit.log
