timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -x -q 2>&1 | tail -3
timeout 200 python tools/bench_attn.py 2>&1 | tail -2 | head -1
