timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --checkpointing 2>&1 | tail -2 | cut -c1-300
