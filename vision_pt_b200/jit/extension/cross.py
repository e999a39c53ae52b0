"""Cross-attention JiT blocks (image tokens attend to context tokens; Lq != Lk) on the sm_100a kernels.

Mirror of /root/reference/src/models/jit/extension/cross.py: `CrossAttention` (32-88) and `CrossJiTBlock` (281-385).
q comes from the image stream with the image RoPE table, k / v from the context stream with the context table; the
reference's `query_mask & key_mask` becomes the per-sample context length (masked QUERY rows are computed like any other:
the reference leaves them undefined and nothing reads them)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...modules.attention import prefix_key_lengths
from ...modules.norm import get_norm_layer
from ..denoiser import Attention, SwiGLU


class CrossAttention(Attention):
    def forward(self, hidden_states, key_value_states, query_cos_sin, key_cos_sin, query_mask=None, key_mask=None):
        B, Lq, D = hidden_states.shape
        Lk = key_value_states.shape[1]
        H, hd = self.num_heads, self.head_dim
        q = self.to_q(hidden_states).view(B, Lq, H, hd)
        k = self.to_k(key_value_states).view(B, Lk, H, hd)
        v = self.to_v(key_value_states).view(B, Lk, H, hd)
        q = ops.qknorm_rope(q, self.q_norm.weight, query_cos_sin, self.q_norm.eps)
        k = ops.qknorm_rope(k, self.k_norm.weight, key_cos_sin, self.k_norm.eps)
        seqlens = prefix_key_lengths(key_mask != 0) if (query_mask is not None and key_mask is not None) else None
        o = ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), seqlens, hd ** -0.5)
        return self.proj_dropout(self.to_o(o.permute(0, 2, 1, 3).reshape(B, Lq, D)))


class CrossJiTBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, mlp_ratio=4.0, attn_dropout=0.0, proj_dropout=0.0, ffn_dropout=0.0, qkv_bias=True,
                 qk_norm=True, bias=True, eps=1e-6, positional_encoding="rope", norm_type="rms", norm_position="sandwich"):
        super().__init__()
        if positional_encoding != "rope":
            raise NotImplementedError("CrossJiTBlock is built for RoPE (PoPE cross attention: PopeAttention.attend)")
        pre = norm_position in ("pre", "sandwich")
        post = norm_position in ("post", "sandwich")
        norm = lambda on: get_norm_layer(norm_type, hidden_dim, eps=eps) if on else nn.Identity()
        self.norm_attn_image_pre, self.norm_attn_post, self.norm_attn_context_pre = norm(pre), norm(post), norm(pre)
        self.attn = CrossAttention(dim=hidden_dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                                   attn_dropout=attn_dropout, proj_dropout=proj_dropout, norm_type="rms")
        self.norm_mlp_pre, self.norm_mlp_post = norm(pre), norm(post)
        self.mlp = SwiGLU(dim=hidden_dim, hidden_dim=int(hidden_dim * mlp_ratio), dropout=ffn_dropout, bias=bias)

    def forward(self, image_hidden_states, context_hidden_states, image_cos_sin, context_cos_sin, image_mask=None,
                context_mask=None):
        h = image_hidden_states + self.norm_attn_post(self.attn(
            self.norm_attn_image_pre(image_hidden_states), self.norm_attn_context_pre(context_hidden_states), image_cos_sin,
            context_cos_sin, query_mask=image_mask, key_mask=context_mask))
        h = h + self.norm_mlp_post(self.mlp(self.norm_mlp_pre(h)))
        return h, context_hidden_states
