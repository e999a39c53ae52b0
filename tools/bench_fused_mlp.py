"""The SwiGLU-fused GEMMs against their unfused composition, CUDA-graph timed with inputs rotated over > L2 of buffers.
  forward : w_2 GEMM + swiglu_fwd kernel          vs   w_2 GEMM with epilogue 1
  backward: w_3 dX GEMM + swiglu_bwd kernel       vs   w_3 dX GEMM with epilogue 2
  python tools/bench_fused_mlp.py        (VPT_LIB=<other build> to compare builds)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import ops  # noqa: E402
from vision_pt_b200.modules.quant import NF4_CODE, nested_code_table  # noqa: E402

dev = torch.device("cuda")
BF = torch.bfloat16
ROT = 3


def rnd(*s, std=1.0):
    return (torch.randn(*s, device=dev) * std).to(BF)


def timed(fn, n=20):
    for i in range(ROT):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        keep = [fn(i % ROT) for i in range(n)]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return 1e3 * e0.elapsed_time(e1) / (5 * n)


print(f"# lib: {os.environ.get('VPT_LIB', 'vision_pt_b200/libvptb200.so')}")
print(f"# {'M, D, F':22s} {'w2+swiglu':>10s} {'fused(1)':>10s} {'plain w2':>10s} | {'w3bwd+swiglu_bwd':>17s} {'fused(2)':>10s} {'plain w3bwd':>12s}   (us)")
for (M, D, F_) in ((21120, 768, 2048), (21120, 1024, 2730), (17600, 1280, 3413)):
    code, ncode = torch.tensor(NF4_CODE), nested_code_table()
    w2 = ops.nf4_quantize(rnd(F_, D, std=0.05), ncode, code)
    w3 = ops.nf4_quantize(rnd(D, F_, std=0.05), ncode, code)
    b2 = rnd(F_, std=0.5)
    d2, u2 = rnd(16, D, std=0.05), rnd(F_, 16, std=0.05)
    d3, u3 = ops._pad_rank(rnd(16, F_, std=0.05), rnd(D, 16, std=0.05))
    hs = [rnd(M, D) for _ in range(ROT)]
    gs = [ops._rows(rnd(M, F_)) for _ in range(ROT)]
    us = [ops._rows(rnd(M, F_)) for _ in range(ROT)]
    dys = [rnd(M, D) for _ in range(ROT)]
    sf = ops.dequant_block([w2, w3], [d2, d3], [u2, u3], transposed=False)
    sb = ops.dequant_block([w2, w3], [d2, d3], [u2, u3], transposed=True)

    def f_plain(i):
        return ops.linear_raw(hs[i], w2, b2, d2, u2, 0.5, want_side=True, scratch=sf[0])

    def f_unfused(i):
        u, t = ops.linear_raw(hs[i], w2, b2, d2, u2, 0.5, want_side=True, scratch=sf[0])
        return ops.swiglu_fwd_raw(gs[i], u), u, t

    def f_fused(i):
        return ops.linear_raw(hs[i], w2, b2, d2, u2, 0.5, gs[i], want_side=True, scratch=sf[0], epilogue=1)

    def b_plain(i):
        return ops.linear_raw(dys[i], w3, None, d3, u3, 0.5, want_side=True, backward=True, scratch=sb[1])

    def b_unfused(i):
        da, t = ops.linear_raw(dys[i], w3, None, d3, u3, 0.5, want_side=True, backward=True, scratch=sb[1])
        return ops.swiglu_bwd_raw(da, gs[i], us[i]), t

    def b_fused(i):
        return ops.linear_raw(dys[i], w3, None, d3, u3, 0.5, gs[i], want_side=True, backward=True, scratch=sb[1], epilogue=2, in2=us[i])

    r = [timed(f) for f in (f_unfused, f_fused, f_plain, b_unfused, b_fused, b_plain)]
    print(f"{M:6d} {D:5d} {F_:5d}     {r[0]:10.1f} {r[1]:10.1f} {r[2]:10.1f} | {r[3]:17.1f} {r[4]:10.1f} {r[5]:12.1f}")
