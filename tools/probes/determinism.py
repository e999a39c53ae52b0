"""Where does run-to-run noise of the LoRA gradients come from?  Runs the same JiT-B step three times and reports, per block
(in backward order), the rel diff of the block-input gradient and of each LoRA gradient between run 0 and run 1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vision_pt_b200 import ops  # noqa: E402
from vision_pt_b200 import train as T  # noqa: E402
from vision_pt_b200.jit import denoiser as dn  # noqa: E402

torch.manual_seed(0)
net = T.build_jit_qlora("JiT-B/16", rank=16, alpha=16.0, device="cuda", seed=42, lora_up_std=0.02)
B, H, W = 4, 256, 256
g = torch.Generator().manual_seed(0)
image = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16).cuda()
t = torch.rand(B, generator=g).to(torch.bfloat16).cuda()
ctx = (torch.randn(B, 64, 768, generator=g) * 0.5).to(torch.bfloat16).cuda()
mask = (torch.arange(64).unsqueeze(0) < torch.tensor([[20], [9], [40], [33]])).to(torch.int64).cuda()
size = torch.tensor([[H, W]]).repeat(B, 1).cuda()
clean = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16).cuda()

# capture dx of every block through the raw ops by wrapping rmsnorm_bwd_raw (its second call per block yields dx)
runs = []
for r in range(3):
    taps = []
    orig_attn, orig_rms, orig_lin = ops.attn_bwd_raw, ops.rmsnorm_bwd_raw, ops.linear_raw

    def attn_bwd(*a, **k):
        out = orig_attn(*a, **k)
        taps.append(("attn_dq", out[0].float().clone()))
        taps.append(("attn_dk", out[1].float().clone()))
        taps.append(("attn_dv", out[2].float().clone()))
        return out

    def rms_bwd(*a, **k):
        out = orig_rms(*a, **k)
        taps.append(("rms_dx", out.float().clone()))
        return out

    def lin(*a, **k):
        out = orig_lin(*a, **k)
        if k.get("backward"):
            taps.append(("lin_bwd", out[0].float().clone()))
            if out[1] is not None:
                taps.append(("lin_bwd_side", out[1][:, :a[0].shape[0]].float().clone()))
        return out

    ops.attn_bwd_raw, ops.rmsnorm_bwd_raw, ops.linear_raw = attn_bwd, rms_bwd, lin
    net.zero_grad(set_to_none=True)
    pred = net(image=image, timestep=t, context=ctx, original_size=size, target_size=size, crop_coords=torch.zeros_like(size),
               context_mask=mask)
    loss = ops.flow_loss(pred, clean, loss_target="image")
    loss.backward()
    torch.cuda.synchronize()
    ops.attn_bwd_raw, ops.rmsnorm_bwd_raw, ops.linear_raw = orig_attn, orig_rms, orig_lin
    grads = {n: p.grad.float().clone() for n, p in net.named_parameters() if p.requires_grad}
    runs.append((pred.float().clone(), taps, grads))

rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
print("pred identical:", torch.equal(runs[0][0], runs[1][0]))
print("\nfirst 60 backward taps (in execution order), rel diff run0 vs run1:")
shown = 0
for (n0, a), (n1, b) in zip(runs[0][1], runs[1][1]):
    d = rel(a, b)
    if shown < 60:
        print(f"  {n0:14s} {tuple(a.shape)!s:22s} {d:.3e}  nan={bool(torch.isnan(a).any())}")
        shown += 1
print("\nLoRA grads, rel diff run0 vs run1 (worst 12) and run0 vs run2:")
rows = sorted(((rel(runs[0][2][n], runs[1][2][n]), rel(runs[0][2][n], runs[2][2][n]), n) for n in runs[0][2]), reverse=True)
for d1, d2, n in rows[:12]:
    print(f"  {d1:.3e} {d2:.3e} {n}")
print("  ... best:", rows[-1])
