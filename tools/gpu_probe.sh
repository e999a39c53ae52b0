timeout 300 python tools/bench_attn.py --iters 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r1e_attnbwd python tools/bench_attn.py --iters 2 > gpurun_out/r1e_ncu.log 2>&1
