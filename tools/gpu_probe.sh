timeout 300 python -m pytest tests/test_gpu_nf4.py -x -q 2>&1 | tail -4
timeout 300 python tools/bench_attn.py --iters 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 -o gpurun_out/r1l_attnfwd python tools/bench_attn.py --iters 2 > gpurun_out/r1l_ncu.log 2>&1
