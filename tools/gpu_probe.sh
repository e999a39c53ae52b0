timeout 400 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | grep -E "^E " | head -20
