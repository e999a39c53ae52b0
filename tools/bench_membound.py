"""HBM-bound kernels at the shapes that matter (north_star part 3: adaLN shift / scale / gate, RMSNorm, patchify, ...):
algorithmic bytes / CUDA-event time against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).

  python tools/bench_membound.py                 # CUDA-graph timed, inputs rotated over > L2 worth of buffers
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/membound.csv python tools/bench_membound.py --once      # DRAM counters per kernel

Shapes: CogView4-class adaLN ([16 x 4096 tokens, 4096]), the JiT-B training shape ([64 x 330, 768]) and the SDXL 640 width;
patchify / unpatchify at the 512-px aspect-ratio buckets of JiT-H (batch 16)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision_pt_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--once", action="store_true", help="one eager pass per case (for ncu), no timing")
args = ap.parse_args()
dev = torch.device("cuda")
BF = torch.bfloat16
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6456.5
ROT = 4          # distinct input sets per case: 4 x (>= 130 MB) > the 126 MB L2 for the big shapes


def rnd(*shape, dtype=BF, std=1.0):
    return (torch.randn(*shape, device=dev) * std).to(dtype)


def timed(fns):
    """fns: one callable per rotation slot.  Returns mean microseconds per call."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    if args.once:
        return None
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fns:
            f()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(3):
            for f in fns:
                f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (5 * 3 * len(fns))


rows_out = []


def case(name, nbytes, fns):
    us = timed(fns)
    if us is None:
        print(f"{name:58s} (one pass)")
        return
    gbs = nbytes / us / 1e3
    rows_out.append((name, nbytes, us, gbs))
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {us:9.1f} us {gbs:8.0f} GB/s  {gbs / PEAK:5.2f}")


def adaln_cases(tag, B, L, D):
    M = B * L
    xs = [rnd(B, L, D) for _ in range(ROT)]
    hs = [rnd(B, L, D) for _ in range(ROT)]
    dys = [rnd(B, L, D) for _ in range(ROT)]
    sc, sh, gate = rnd(B, D, std=0.3), rnd(B, D, std=0.3), rnd(B, D, std=0.3)
    y = torch.empty_like(xs[0])
    mean = torch.empty(M, dtype=torch.float32, device=dev)
    rstd = torch.empty(M, dtype=torch.float32, device=dev)
    call = ops._lib.call
    p, st = ops._p, ops._stream
    case(f"ln_modulate_fwd   {tag}", 4 * M * D, [lambda x=x: call("vpt_ln_modulate_fwd", p(x), p(sc), p(sh), p(y), p(mean), p(rstd), M, L, D, 1e-5, st()) for x in xs])
    case(f"ln_modulate_bwd   {tag} (frozen modulation)", 6 * M * D, [lambda x=x, d=d: call("vpt_ln_modulate_bwd", p(d), p(x), p(sc), p(mean), p(rstd), p(y), None, None, M, L, D, st()) for x, d in zip(xs, dys)])
    case(f"gate_residual_fwd {tag}", 6 * M * D, [lambda x=x, h=h: call("vpt_gate_residual_fwd", p(x), p(h), p(gate), p(y), M, L, D, st()) for x, h in zip(xs, hs)])
    case(f"gate_residual_bwd {tag} (frozen gate)", 6 * M * D, [lambda d=d, h=h: call("vpt_gate_residual_bwd", p(d), p(h), p(gate), p(y), None, M, L, D, st()) for d, h in zip(dys, hs)])
    w, b = rnd(D), rnd(D)
    case(f"layernorm_fwd     {tag} (affine)", 4 * M * D, [lambda x=x: call("vpt_layernorm_fwd", p(x), p(w), p(b), p(y), p(mean), p(rstd), M, D, 1e-5, st()) for x in xs])
    case(f"layernorm_bwd     {tag} (frozen affine)", 6 * M * D, [lambda x=x, d=d: call("vpt_layernorm_bwd", p(d), p(x), p(w), p(mean), p(rstd), p(y), None, None, M, D, st()) for x, d in zip(xs, dys)])
    if D <= 2048:
        ys = [ops.rmsnorm_fwd_raw(x.view(M, D), w, 1e-6)[1] for x in xs[:1]]
        case(f"rmsnorm_fwd       {tag}", 4 * M * D, [lambda x=x: ops.rmsnorm_fwd_raw(x.view(M, D), w, 1e-6) for x in xs])
        case(f"rmsnorm_bwd       {tag} (+ residual grad)", 8 * M * D, [lambda x=x, d=d, h=h: ops.rmsnorm_bwd_raw(d.view(M, D), x.view(M, D), w, ys[0], h.view(M, D), 1e-6) for x, d, h in zip(xs, dys, hs)])


print(f"# HBM-bound kernels: algorithmic bytes / CUDA-event time; peak = {PEAK} GB/s (MEASURED_PEAKS.json hbm_gbs)")
print(f"# {'kernel / shape':56s} {'bytes':>12s} {'time':>12s} {'rate':>13s}  frac")
adaln_cases("CogView4-class [16 x 4096, 4096]", 16, 4096, 4096)
adaln_cases("JiT-B train    [64 x 330, 768]", 64, 330, 768)
adaln_cases("SDXL 640       [4 x 4096, 640]", 4, 4096, 640)

# gated activations (GeGLU at the SDXL 640 width: F = 2560; SwiGLU at JiT-B: F = 2048)
for tag, M, F_, kind in (("GeGLU  SDXL 640 [16384, 2560]", 16384, 2560, 1), ("SwiGLU JiT-B    [21120, 2048]", 21120, 2048, 0)):
    proj = [rnd(M, 2 * F_) for _ in range(ROT)]
    das = [rnd(M, F_) for _ in range(ROT)]
    a = torch.empty(M, F_, dtype=BF, device=dev)
    dh, dg = torch.empty_like(a), torch.empty_like(a)
    call, p, st = ops._lib.call, ops._p, ops._stream
    case(f"gated_act_fwd     {tag}", 6 * M * F_, [lambda t=t: call("vpt_gated_act_fwd", p(t[:, :F_]), p(t[:, F_:]), p(a), M, F_, 2 * F_, 2 * F_, F_, kind, st()) for t in proj])
    case(f"gated_act_bwd     {tag}", 10 * M * F_, [lambda t=t, d=d: call("vpt_gated_act_bwd", p(d), p(t[:, :F_]), p(t[:, F_:]), p(dh), p(dg), M, F_, F_, 2 * F_, 2 * F_, F_, F_, kind, st()) for t, d in zip(proj, das)])

# patchify / unpatchify at the 512-px buckets (JiT-H, batch 16) and the JiT-B batch
for (B, H, W) in ((16, 512, 512), (16, 448, 576), (16, 256, 1024), (64, 256, 256)):
    imgs = [rnd(B, 3, H, W) for _ in range(ROT)]
    n = B * 3 * H * W
    case(f"patchify   (c,py,px)  [{B}, 3, {H}, {W}]", 4 * n, [lambda t=t: ops.patchify_op(t, 16, 0) for t in imgs])
    pts = [rnd(B, (H // 16) * (W // 16), 768) for _ in range(ROT)]
    case(f"unpatchify (py,px,c)  [{B}, 3, {H}, {W}]", 4 * n, [lambda t=t: ops.unpatchify_op(t, 3, H, W, 16, 1) for t in pts])

# token gather (TREAD) and the half-split rotary embedding
x3 = [rnd(64, 256, 768) for _ in range(ROT)]
idx = torch.randperm(256, device=dev)[:128].contiguous()
case("token_gather     [64, 256 -> 128, 768]", 4 * 64 * 128 * 768, [lambda t=t: ops.token_gather(t, idx) for t in x3])
q4 = [rnd(2, 4160, 32, 128) for _ in range(ROT)]
cos, sin = torch.randn(4096, 128, device=dev), torch.randn(4096, 128, device=dev)
case("rope_half        [2, 4160, 32, 128] (64 text tokens)", 4 * q4[0].numel(), [lambda t=t: ops.rope_half(t, cos, sin, 64) for t in q4])
# the step's glue (JiT-B batch): noise preparation, the patch-token slice and its gradient, the loss's upstream gradient
ims = [rnd(64, 3, 256, 256, dtype=torch.float16) for _ in range(ROT)]
zs = [rnd(64, 3, 256, 256, dtype=torch.float16) for _ in range(ROT)]
ts = torch.rand(64, device=dev)
n_img = ims[0].numel()
case("noise_mix        fp16 [64, 3, 256, 256] (+ bf16 copy)", 8 * n_img, [lambda a=a, z=z: ops.noise_mix(a, z, ts, 1.0) for a, z in zip(ims, zs)])
tok = [rnd(64, 330, 768) for _ in range(ROT)]
case("token_prefix     [64, 330 -> 256, 768]", 4 * 64 * 256 * 768, [lambda t=t: ops.packed_tokens(t[:, :256]) for t in tok])
dps = [rnd(64, 3, 256, 256) for _ in range(ROT)]
one = torch.ones(1, device=dev)
outs = [torch.empty_like(d) for d in dps]
case("scale_by_scalar  bf16 [64, 3, 256, 256]", 4 * n_img,
     [lambda d=d, o=o: ops._lib.call("vpt_scale_by_scalar", ops._p(d), ops._p(one), ops._p(o), d.numel(), ops._stream()) for d, o in zip(dps, outs)])
if rows_out:
    json.dump([{"kernel": n, "bytes": b, "us": u, "gbs": g, "frac": g / PEAK} for n, b, u, g in rows_out],
              open(os.path.join(ROOT, "gpurun_out", "membound.json"), "w"), indent=1)
