timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --model JiT-H/16 --res 512 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1q_jith.json; python -c "
import json; d=json.load(open('gpurun_out/r1q_jith.json')); print('JiT-H', {k:d[k] for k in ('value','ms_per_step','achieved_tflops_step')}, d['roofline']['frac'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('JiT-B', {k:d[k] for k in ('value','ms_per_step')})"
