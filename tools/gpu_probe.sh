timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_elementwise.py tests/test_gpu_block.py -x -q 2>&1 | tail -12
