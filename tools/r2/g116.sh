cd /root/repo
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -x -q -k "three_training_steps or uncleared or golden or use_dc_network" > gpurun_out/r2_sanitizer_model.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_sanitizer_model.log
