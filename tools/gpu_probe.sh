timeout 600 python -m pytest tests/test_gpu_linear.py -x -q -k sdxl 2>&1 | tail -5
