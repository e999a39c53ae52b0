timeout 200 python tools/bench_linear.py --model L --no-check --no-res 2>&1 | tail -6 | cut -c1-75
timeout 200 python tools/bench_linear.py --model L --no-check --no-res --tile 128 2>&1 | tail -6 | cut -c1-75
timeout 300 python bench.py --model JiT-L/16 --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','achieved_tflops_step')}); print(d['roofline']['achieved'], d['roofline']['frac'])"
