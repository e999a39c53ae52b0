timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_attn.py 2>&1 | tail -3
timeout 300 python tools/bench_attn.py --B 16 --H 16 --L 1100 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
