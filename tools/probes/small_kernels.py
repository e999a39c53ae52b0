"""Launches the small HBM-bound kernels once each on JiT-B step shapes (for an ncu duration pass)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vision_pt_b200 import ops  # noqa: E402

B, C, H, W = 64, 3, 256, 256
pred = torch.randn(B, C, H, W, device="cuda").to(torch.bfloat16).requires_grad_(True)
clean = torch.randn(B, C, H, W, device="cuda").to(torch.float16)
for _ in range(3):
    ops.flow_loss(pred, clean).backward()
Hh, L = 12, 330
mk = lambda: torch.randn(B, L, Hh, 64, device="cuda").to(torch.bfloat16).permute(0, 2, 1, 3)
q, k, v, d_o = mk(), mk(), mk(), mk()
o, lse = ops.attn_fwd_raw(q, k, v, None, 0.125)
for _ in range(3):
    ops.attn_bwd_raw(q, k, v, o, d_o, lse, None, 0.125)
torch.cuda.synchronize()
