"""Times the attention forward / backward kernels (CUDA graph of `iters` calls).  python tools/bench_attn.py [--B 64 --H 12 --L 330]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--H", type=int, default=12)
ap.add_argument("--L", type=int, default=330)
ap.add_argument("--Lk", type=int, default=0, help="key length (0 = L): cross-attention")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--hd", type=int, default=64)
args = ap.parse_args()
B, H, L = args.B, args.H, args.L
Lk = args.Lk or L
dev = torch.device("cuda")
torch.manual_seed(0)
HD = args.hd
mk = lambda n=L: torch.randn(B, n, H, HD, device=dev).to(torch.bfloat16).permute(0, 2, 1, 3)
sets = [(mk(), mk(Lk), mk(Lk), mk()) for _ in range(4)]
seq = torch.randint(max(1, Lk - 56), Lk + 1, (B,), device=dev, dtype=torch.int32)
fl_f = 4.0 * B * H * L * Lk * HD


def timed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(2):
            fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [fn(i) for i in range(args.iters)]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 1e3 * best / args.iters


fw = lambda i: ops.attn_fwd_raw(sets[i % 4][0], sets[i % 4][1], sets[i % 4][2], seq, HD ** -0.5)
us = timed(fw)
print(f"fwd  B={B} H={H} L={L}x{Lk}: {us:8.1f} us  {fl_f / us / 1e6:7.1f} TF (full LxL)")
outs = [ops.attn_fwd_raw(s[0], s[1], s[2], seq, HD ** -0.5) for s in sets]
bw = lambda i: ops.attn_bwd_raw(sets[i % 4][0], sets[i % 4][1], sets[i % 4][2], outs[i % 4][0], sets[i % 4][3], outs[i % 4][1], seq, HD ** -0.5)
us = timed(bw)
print(f"bwd  B={B} H={H} L={L}x{Lk}: {us:8.1f} us  {2.5 * fl_f / us / 1e6:7.1f} TF (incl. dq zero-fill + delta pre-pass)")
