"""U-JiT block (long skip connections merged by Linear(2D -> D), sandwich norms) on the sm_100a kernels.

Mirror of /root/reference/src/models/jit/extension/uvit.py:28-147 (`UJiTBlock`); module names are the reference's
(`skip_merge`, `norm_attn_pre/post`, `attn`, `norm_mlp_pre/post`, `mlp`).  Attention / SwiGLU / norms are the JiT block's
kernels; the skip merge is one more fused (NF4 + LoRA capable) linear over the concatenated features."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...modules.norm import get_norm_layer
from ..denoiser import Attention, SwiGLU
from .pope import PopeAttention


class UJiTBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, mlp_ratio=4.0, attn_dropout=0.0, proj_dropout=0.0, ffn_dropout=0.0, qkv_bias=True,
                 qk_norm=True, bias=True, has_skip_connection=False, eps=1e-6, positional_encoding="rope", norm_type="rms",
                 norm_position="sandwich"):
        super().__init__()
        self.has_pre_norm = norm_position in ("pre", "sandwich")
        self.has_post_norm = norm_position in ("post", "sandwich")
        if has_skip_connection:
            self.skip_merge = nn.Linear(hidden_dim * 2, hidden_dim, bias=bias)
        norm = lambda on: get_norm_layer(norm_type, hidden_dim, eps=eps) if on else nn.Identity()
        self.norm_attn_pre, self.norm_attn_post = norm(self.has_pre_norm), norm(self.has_post_norm)
        attn_cls = PopeAttention if positional_encoding in ("pope", "n-pope") else Attention
        self.attn = attn_cls(dim=hidden_dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm, attn_dropout=attn_dropout,
                             proj_dropout=proj_dropout, norm_type="rms")
        self.norm_mlp_pre, self.norm_mlp_post = norm(self.has_pre_norm), norm(self.has_post_norm)
        self.mlp = SwiGLU(dim=hidden_dim, hidden_dim=int(hidden_dim * mlp_ratio), dropout=ffn_dropout, bias=bias)

    def forward(self, hidden_states, cos_sin, skip_hidden_states=None, seqlens=None):
        if skip_hidden_states is not None:
            hidden_states = self.skip_merge(torch.cat([hidden_states, skip_hidden_states], dim=-1))
        hidden_states = hidden_states + self.norm_attn_post(self.attn(self.norm_attn_pre(hidden_states), cos_sin, seqlens))
        hidden_states = hidden_states + self.norm_mlp_post(self.mlp(self.norm_mlp_pre(hidden_states)))
        return hidden_states
