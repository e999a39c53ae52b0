timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_train.py tests/test_abi.py -x -q 2>&1 | tail -15
