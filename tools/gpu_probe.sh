timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 500 python bench.py --model JiT-H/16 --res 512 --batch 16 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
