mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_fullsize.py tests/test_gpu_train.py tests/test_gpu_trainer.py tests/test_gpu_linear.py -q -x > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^E  |Error|FAILED" gpurun_out/r2p_tests.log | head -12
timeout 500 python tools/ab_step.py > gpurun_out/r2p_ab_jitb.txt 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2p_ab_jitb.txt
