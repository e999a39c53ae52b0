"""PoPE positional encoding (softplus magnitude x position phase) on the fused kernel.

Mirror of /root/reference/src/models/jit/extension/pope.py: `apply_pope` (6-38), `PopeEmbedder` table math (41-190), and
`PopeAttention` of /root/reference/src/models/jit/denoiser.py:398-474.  The position phases are batch-invariant, so they
are kept as one fp32 (cos, sin) table [L, d, 2] per bucket; softplus, the learned per-head phase bias and the rotation
are one kernel (`vpt_pope_fwd` / `_bwd`).  PoPE doubles the q / k width (re, im) while v keeps head_dim: the score GEMM
runs at 2 * head_dim and v is zero-padded to that width for the attention kernel (an experimental variant of the
reference; correct first, not tuned).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops


class PopeEmbedder:
    """Position-phase tables: per axis, angle = position * theta^(-k/dim) for k in [0, dim) (PoPE uses the full dim)."""

    def __init__(self, pope_theta: float = 256.0, axes_dims=(64, 128, 128), axes_lens=(256, 128, 128),
                 zero_centered=(False, True, True)):
        self.pope_theta, self.axes_dims, self.axes_lens = pope_theta, list(axes_dims), list(axes_lens)
        self.zero_centered = list(zero_centered)
        self.num_axes = len(self.axes_dims)

    def angles(self, position_ids: torch.Tensor) -> torch.Tensor:
        """position_ids [L, n_axes] -> fp32 angles [L, sum(axes_dims)] (same numbers as the reference's gathered tables:
        float64 outer product rounded to fp32)."""
        cols = []
        for i, dim in enumerate(self.axes_dims):
            freqs = 1.0 / (self.pope_theta ** (torch.arange(0, dim, 1, dtype=torch.float64) / dim))
            cols.append(torch.outer(position_ids[:, i].to(torch.float64), freqs).float())
        return torch.cat(cols, dim=-1)

    def prepare_image_position_ids(self, height: int, width: int, patch_size: int, global_index: int) -> torch.Tensor:
        hp, wp = height // patch_size, width // patch_size
        pos = torch.zeros(hp, wp, self.num_axes)
        pos[:, :, 0] = global_index
        pos[:, :, 1] = torch.arange(hp // 2 - hp, hp // 2).unsqueeze(1).repeat(1, wp)
        pos[:, :, 2] = torch.arange(wp // 2 - wp, wp // 2).unsqueeze(0).repeat(hp, 1)
        return pos.view(-1, self.num_axes)

    def prepare_context_position_ids(self, seq_len: int, global_index: int = 0) -> torch.Tensor:
        pos = torch.zeros(seq_len, self.num_axes)
        pos[:, 0] = global_index
        pos[:, 1] = torch.arange(seq_len)
        pos[:, 2] = torch.arange(seq_len)
        return pos

    def __call__(self, position_ids: torch.Tensor) -> torch.Tensor:
        return pope_table(self.angles(position_ids))


def pope_table(angles_or_cis: torch.Tensor) -> torch.Tensor:
    """fp32 [L, d, 2] (cos, sin) table from fp32 angles [L, d] or from the reference's complex freqs_cis [L, d]."""
    if angles_or_cis.is_complex():
        return torch.stack([angles_or_cis.real, angles_or_cis.imag], dim=-1).float().contiguous()
    cis = torch.polar(torch.ones_like(angles_or_cis), angles_or_cis)          # as the reference builds it
    return torch.stack([cis.real, cis.imag], dim=-1).contiguous()


def apply_pope(inputs: torch.Tensor, cos_sin: torch.Tensor, learned_bias: torch.Tensor | None = None) -> torch.Tensor:
    """Reference signature: inputs [B, H, L, d] -> [B, H, L, 2d]; `cos_sin` = pope_table(...) [L, d, 2]."""
    y = ops.pope(inputs.permute(0, 2, 1, 3), cos_sin, learned_bias)           # token-major inside
    return y.permute(0, 2, 1, 3)


class PopeAttention(nn.Module):
    """Attention with PoPE on q and k (learned phase bias on k only, clamped to [-pi, pi]); names as the reference's."""

    def __init__(self, dim, num_heads=8, qkv_bias=True, qk_norm=True, attn_dropout=0.0, proj_dropout=0.0, eps=1e-6,
                 norm_type="rms"):
        super().__init__()
        from ...modules.norm import get_norm_layer
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.q_norm = get_norm_layer(norm_type, self.head_dim, eps=eps) if qk_norm else nn.Identity()
        self.k_norm = get_norm_layer(norm_type, self.head_dim, eps=eps) if qk_norm else nn.Identity()
        self.to_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_k = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_v = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.to_o = nn.Linear(dim, dim)
        self.proj_dropout = nn.Dropout(proj_dropout)
        self.pope_bias = nn.Parameter(torch.zeros((num_heads, self.head_dim)))

    def attend(self, q4, k4, v4, q_table, k_table, seqlens):
        """q4 / k4 / v4 token-major [B, L, H, hd] (already QK-normed)."""
        H, hd = self.num_heads, self.head_dim
        q = ops.pope(q4, q_table, None)
        k = ops.pope(k4, k_table, self.pope_bias.detach().clamp(-math.pi, math.pi))
        v = torch.nn.functional.pad(v4, (0, hd))                      # [.., 2hd]: the kernel wants one head width
        o = ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), seqlens, (2 * hd) ** -0.5)
        return o.permute(0, 2, 1, 3)[..., :hd]

    def forward(self, hidden_states, cos_sin, seqlens=None):
        B, L, D = hidden_states.shape
        H, hd = self.num_heads, self.head_dim
        q = self.q_norm(self.to_q(hidden_states).view(B, L, H, hd))
        k = self.k_norm(self.to_k(hidden_states).view(B, L, H, hd))
        v = self.to_v(hidden_states).view(B, L, H, hd)
        o = self.attend(q, k, v, cos_sin, cos_sin, seqlens)
        return self.proj_dropout(self.to_o(o.reshape(B, L, D)))
