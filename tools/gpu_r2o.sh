mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_blocks_ext.py -q 2>&1 | tail -4
timeout 300 python tools/bench_membound.py 2>&1 | grep -E "ln_modulate|layernorm"
