timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_block.py tests/test_gpu_train.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['clocks'])"
