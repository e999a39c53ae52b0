timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_attn.py 2>&1 | tail -3
timeout 300 python tools/bench_attn.py --B 16 --H 16 --L 1100 2>&1 | tail -3
