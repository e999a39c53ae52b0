mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_linear.py -q -x -k "swiglu_epilogues" 2>&1 | tail -2
timeout 300 python tools/bench_fused_mlp.py 2>&1 | tail -5
VPT_LIB=build/libvptb200_m2s4.so timeout 300 python -m pytest tests/test_gpu_linear.py -q -x -k "swiglu_epilogues" 2>&1 | tail -2
VPT_LIB=build/libvptb200_m2s4.so timeout 300 python tools/bench_fused_mlp.py 2>&1 | tail -5
