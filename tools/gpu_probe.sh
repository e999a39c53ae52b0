timeout 900 python -m pytest tests/test_gpu_linear.py -x -q -k "full_size" 2>&1 | tail -5
