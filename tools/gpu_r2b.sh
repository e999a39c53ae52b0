mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_block.py -q -s > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
grep -n "^\[\|checkpointing:\|passed\|failed" gpurun_out/r2b_tests.log | tail -12; cat gpurun_out/r2b_bench.json | head -c 4000; tail -5 gpurun_out/r2b_bench.err
