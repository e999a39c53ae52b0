# Short probe on one B200: the tests of what changed, the non-library kernels of one step, the step's launch list.
#   gpurun --timeout 900 -- 'bash tools/gpu_probe.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_elementwise.py tests/test_gpu_train.py tests/test_gpu_block.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/probe_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^E  |FAILED" gpurun_out/probe_tests.log | head -10
timeout 300 python tools/profile_aten.py > gpurun_out/probe_aten.txt 2>&1; echo "aten rc=$?"
python tools/profile_step.py > gpurun_out/probe_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/probe_step_launches.csv python tools/profile_step.py > gpurun_out/probe_launches.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/probe_step_launches.csv > gpurun_out/probe_step_launches_summary.txt 2>&1; head -45 gpurun_out/probe_step_launches_summary.txt
