timeout 400 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -x -q 2>&1 | tail -2
timeout 200 python tools/bench_attn.py 2>&1 | tail -2 | head -1
timeout 200 python tools/bench_attn.py --L 1100 --B 16 2>&1 | tail -2 | head -1
