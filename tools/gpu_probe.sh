timeout 900 python tools/bench_arb.py --steps 36 2>&1 | tail -1 > gpurun_out/r1q_arb.json; cat gpurun_out/r1q_arb.json | cut -c1-1500
