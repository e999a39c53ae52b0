timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -x -q 2>&1 | tail -3
VPT_ATTN_PROF=1 timeout 300 python tools/bench_attn.py 2>&1 | grep -E "vpt|bwd" | head -4
timeout 300 python tools/bench_attn.py --B 16 --H 16 --L 1100 2>&1 | tail -2
