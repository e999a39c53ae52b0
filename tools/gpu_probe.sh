timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1o_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r1o_bench.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
timeout 400 python bench.py --model JiT-L/16 --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('JiT-L', {k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --checkpointing 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ckpt', {k:d[k] for k in ('value','ms_per_step')})"
