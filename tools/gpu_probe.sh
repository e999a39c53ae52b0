timeout 300 python -m pytest tests/test_gpu_attention.py -x -q 2>&1 | tail -2
VPT_ATTN_PROF=1 timeout 300 python tools/bench_attn.py 2>&1 | grep -E "vpt|fwd" | head -4
