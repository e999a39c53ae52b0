mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blocks_ext.py tests/test_gpu_elementwise.py -q > gpurun_out/r2h_tests.log 2>&1; echo "ext tests rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/r2h_tests.log | head -30
timeout 600 python tools/bench_membound.py > gpurun_out/r2h_membound.txt 2>&1; echo "membound rc=$?"
cat gpurun_out/r2h_membound.txt | tail -45
