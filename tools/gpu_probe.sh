timeout 200 python tools/bench_linear.py --model B --no-check --no-res 2>&1 | tail -6
timeout 200 python tools/bench_linear.py --model L --no-check --no-res 2>&1 | tail -6
