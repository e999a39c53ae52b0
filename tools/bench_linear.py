"""Times the fused NF4-LoRA linear (forward and backward-dX calls, dequantisation included) at the block shapes and
checks each result against torch on the dequantised weight.  python tools/bench_linear.py [--model B|L|H] [--tile 0|128|192]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import ops  # noqa: E402
from vision_pt_b200.modules.quant import nested_code_table  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="B", help="B | L | H (JiT block linears) | sdxl640 | sdxl1280 (SDXL TransformerBlock linears, SURVEY 8a12)")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--M", type=int, default=21120)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--no-res", action="store_true")
ap.add_argument("--only", type=int, default=-1)
args = ap.parse_args()
if args.model.startswith("sdxl"):
    D = int(args.model[4:])
    # attn1/attn2 to_q/k/v/out [D,D], attn2.to_k/v [D,2048] (ctx), ff.net.0.proj D->8D (GeGLU), ff.net.2 4D->D
    shapes = [(D, D), (2048, D), (D, 8 * D), (4 * D, D)]
else:
    D, F = {"B": (768, 2048), "L": (1024, 2730), "H": (1280, 3413)}[args.model]
    shapes = [(D, D), (D, F), (F, D)]           # (K, N)
dev = torch.device("cuda")
torch.manual_seed(0)
code = torch.tensor([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                     -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                     0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                     0.7229568362236023, 1.0])
M = args.M
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for si, (K, N) in enumerate(shapes):
    if args.only >= 0 and si != args.only:
        continue
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    n = N * K
    pad = (-n) % 64
    wq = torch.cat([w.reshape(-1), w.new_zeros(pad)]).view(-1, 1) if pad else w
    qs = ops.nf4_quantize(wq.reshape(-1, 64) if pad else w, nested_code_table(), code)
    qs.shape = (N, K)
    wd = ops.nf4_dequantize(ops.Nf4Tensors(qs.packed, qs.absmax, qs.nested_absmax, qs.nested_code, qs.code, qs.offset,
                                           ((n + pad) // 64, 64), torch.bfloat16)).reshape(-1)[:n].view(N, K)
    down = (torch.randn(16, K, device=dev) * 0.05).to(torch.bfloat16)
    up = (torch.randn(N, 16, device=dev) * 0.05).to(torch.bfloat16)
    dpad, upad = ops._pad_rank(down, up)
    bias = (torch.randn(N, device=dev) * 0.1).to(torch.bfloat16)
    for bwd in (False, True):
        cin = N if bwd else K
        cout = K if bwd else N
        ld = (cin + 7) // 8 * 8
        x = torch.randn(M, ld, device=dev).to(torch.bfloat16)[:, :cin]
        res = torch.randn(M, (cout + 7) // 8 * 8, device=dev).to(torch.bfloat16)[:, :cout]
        call = lambda: ops.linear_raw(x, qs, None if bwd else bias, dpad, upad, 1.0, res, want_side=True, backward=bwd,
                                      tile_n=args.tile)
        y, side = call()
        torch.cuda.synchronize()
        err = float("nan")
        if not args.no_check:
            xs = x[:4096].float()
            if not bwd:
                t = (xs @ down.float().t()).to(torch.bfloat16).float()
                ref = xs @ wd.float().t() + bias.float() + t @ up.float().t() + res[:4096].float()
            else:
                t = (xs @ up.float()).to(torch.bfloat16).float()
                ref = xs @ wd.float() + t @ down.float() + res[:4096].float()
            err = float((y[:4096].float() - ref).abs().max() / ref.abs().max())
            serr = float((side[:, :4096].t().float() - t).abs().max() / t.abs().max().clamp_min(1e-9))
            tail = float((y[-1].float() - (
                (x[-1:].float() @ (wd.float().t() if not bwd else wd.float())) + (0 if bwd else bias.float())
                + ((x[-1:].float() @ (down.float().t() if not bwd else up.float())).to(torch.bfloat16).float()
                   @ (up.float().t() if not bwd else down.float())) + res[-1:].float())[0]).abs().max())
        # timing: one CUDA graph of `iters` calls that rotate over input/output sets larger than the 126 MB L2
        # (no CPU launch overhead in the measurement, no L2-warm repeats)
        nset = 1 if args.no_flush else max(2, int(300e6 // (M * (cin + 2 * cout) * 2)) + 1)
        xs_ = [x] + [torch.randn(M, ld, device=dev).to(torch.bfloat16)[:, :cin] for _ in range(nset - 1)]
        rs_ = [res] + [torch.randn(M, (cout + 7) // 8 * 8, device=dev).to(torch.bfloat16)[:, :cout] for _ in range(nset - 1)]
        def run(i):
            return ops.linear_raw(xs_[i % nset], qs, None if bwd else bias, dpad, upad, 1.0, None if args.no_res else rs_[i % nset], want_side=True,
                                  backward=bwd, tile_n=args.tile)
        side_stream = torch.cuda.Stream()
        with torch.cuda.stream(side_stream):
            for i in range(3):
                run(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = [run(i) for i in range(args.iters)]
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 1e30
        for _ in range(3):
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            tot = min(tot, e0.elapsed_time(e1))
        del keep, graph
        us = 1e3 * tot / args.iters
        # vendor reference on the same shape: torch.matmul (cuBLAS) bf16, no LoRA / bias / residual, weight already dense
        wden = torch.randn(cout, cin, device=dev).to(torch.bfloat16)
        xs_c = [t.contiguous() for t in xs_]
        g2 = torch.cuda.CUDAGraph()
        for i in range(2):
            torch.matmul(xs_c[i % nset], wden.t())
        torch.cuda.synchronize()
        with torch.cuda.graph(g2):
            keep2 = [torch.matmul(xs_c[i % nset], wden.t()) for i in range(args.iters)]
        g2.replay(); torch.cuda.synchronize()
        best2 = 1e30
        for _ in range(3):
            e0.record(); g2.replay(); e1.record(); torch.cuda.synchronize()
            best2 = min(best2, e0.elapsed_time(e1))
        us_cublas = 1e3 * best2 / args.iters
        del keep2, g2
        fl = 2.0 * M * K * N + 2.0 * M * 16 * (K + N)
        print(f"K={K:5d} N={N:5d} {'bwd' if bwd else 'fwd'} tile={args.tile:3d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF   [cuBLAS plain GEMM {us_cublas:6.1f} us {2.0 * M * K * N / us_cublas / 1e6:7.1f} TF]  "
              f"rel err {err:.2e} side {serr if not args.no_check else float('nan'):.2e} tail {tail if not args.no_check else float('nan'):.2e}", flush=True)
