"""SDXL UNet transformer block on the sm_100a kernels (BASELINE.json configs[4]).

Mirror of /root/reference/src/models/sdxl/denoiser.py: SelfAttention (32-94), CrossAttention (97-172), GeGLU (175-186),
FeedForward (189-207), TransformerBlock (213-280).  Module and parameter names are the reference's (`attn1.to_q`,
`attn2.to_k`, `ff.net.0.proj`, `ff.net.2`, `norm1..3`), so its LoRA / quantisation key lists (`attn1`, `attn2`, `.ff.`:
configs/sdxl/flow_match/config.yml:19-28) select the same linears.  Every linear is an nn.Linear that the quant / PEFT
registries swap for NF4Linear / LoRALinear (fused tcgen05 GEMM); attention runs the tcgen05 kernel on token-major q / k / v
(no [B,H,L,hd] copies; cross attention has Lq != Lk); the three affine LayerNorms and h * gelu(gate) are single fused
passes.  Only the transformer block is built -- the convolutional UNet around it is out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..modules.attention import AttentionImplementation, key_lengths_from_mask


def _attend(q, k, v, num_heads: int, head_dim: int, mask=None):
    """q [B, Lq, H*hd], k / v [B, Lk, H*hd] -> [B, Lq, H*hd]; heads are strided views, nothing is transposed in memory."""
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    as4 = lambda t, L: t.reshape(B, L, num_heads, head_dim).permute(0, 2, 1, 3)
    seqlens = key_lengths_from_mask(mask, B, Lk)
    o = ops.attention(as4(q, Lq), as4(k, Lk), as4(v, Lk), seqlens, head_dim ** -0.5)
    return o.permute(0, 2, 1, 3).reshape(B, Lq, num_heads * head_dim)


class SelfAttention(nn.Module):
    def __init__(self, num_heads: int, head_dim: int, dropout: float, attn_implementation: AttentionImplementation = "eager"):
        super().__init__()
        self.inner_dim = num_heads * head_dim
        self.num_heads, self.head_dim, self.dropout = num_heads, head_dim, dropout
        self.attn_implementation = attn_implementation
        self.to_q = nn.Linear(self.inner_dim, self.inner_dim, bias=False)
        self.to_k = nn.Linear(self.inner_dim, self.inner_dim, bias=False)
        self.to_v = nn.Linear(self.inner_dim, self.inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(self.inner_dim, self.inner_dim), nn.Dropout(dropout))

    def forward(self, hidden_states: torch.Tensor, mask: torch.Tensor | None = None):
        q, k, v = self.to_q(hidden_states), self.to_k(hidden_states), self.to_v(hidden_states)
        return self.to_out(_attend(q, k, v, self.num_heads, self.head_dim, mask))


class CrossAttention(nn.Module):
    def __init__(self, query_dim: int, context_dim: int, num_heads: int, head_dim: int, dropout: float,
                 attn_implementation: AttentionImplementation = "eager"):
        super().__init__()
        self.query_dim, self.context_dim = query_dim, context_dim
        self.inner_dim = num_heads * head_dim
        self.num_heads, self.head_dim, self.dropout = num_heads, head_dim, dropout
        self.attn_implementation = attn_implementation
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, self.inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, self.inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(self.inner_dim, query_dim), nn.Dropout(dropout))

    def forward(self, query: torch.Tensor, context: torch.Tensor, mask: torch.Tensor | None = None,
                time_embedding: torch.Tensor | None = None, *args, **kwargs) -> torch.Tensor:
        q, k, v = self.to_q(query), self.to_k(context), self.to_v(context)
        return self.to_out(_attend(q, k, v, self.num_heads, self.head_dim, mask))


class GeGLU(nn.Module):
    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.proj = nn.Linear(in_dim, out_dim * 2)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        hidden, gate = self.proj(hidden_states).chunk(2, dim=-1)     # two strided views of one GEMM output
        return ops.gated_act(hidden, gate, "gelu")                   # hidden * gelu(gate), one pass


class FeedForward(nn.Module):
    def __init__(self, hidden_dim: int, multiplier: float = 4, dropout: float = 0.0):
        super().__init__()
        self.intermediate_dim = int(hidden_dim * multiplier)
        self.net = nn.Sequential(GeGLU(hidden_dim, self.intermediate_dim), nn.Dropout(dropout),
                                 nn.Linear(self.intermediate_dim, hidden_dim, bias=True))

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        return self.net(hidden_states)


class _LayerNorm(nn.LayerNorm):
    """nn.LayerNorm (affine, eps 1e-5) as one fused pass."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(x, self.weight, self.bias, self.eps)


class TransformerBlock(nn.Module):
    self_attention_class: type[SelfAttention] = SelfAttention
    cross_attention_class: type[CrossAttention] = CrossAttention

    def __init__(self, hidden_dim: int, num_heads: int, head_dim: int, context_dim: int = 2048,
                 attn_implementation: AttentionImplementation = "eager"):
        super().__init__()
        self.hidden_dim, self.num_heads, self.head_dim, self.context_dim = hidden_dim, num_heads, head_dim, context_dim
        self.attn_implementation = attn_implementation
        self.attn1 = self.self_attention_class(num_heads=num_heads, head_dim=head_dim, dropout=0.0,
                                               attn_implementation=attn_implementation)
        self.ff = FeedForward(hidden_dim=hidden_dim, dropout=0.0)
        self.attn2 = self.cross_attention_class(query_dim=hidden_dim, context_dim=context_dim, num_heads=num_heads,
                                                head_dim=head_dim, dropout=0.0, attn_implementation=attn_implementation)
        self.norm1 = _LayerNorm(hidden_dim)
        self.norm2 = _LayerNorm(hidden_dim)
        self.norm3 = _LayerNorm(hidden_dim)

    @staticmethod
    def _residual(linear_seq: nn.Sequential, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """residual + to_out(x): folded into the output GEMM's epilogue when that layer takes a residual."""
        lin = linear_seq[0]
        from ..modules.peft import LoRALinear
        if isinstance(lin, LoRALinear) and x.is_cuda:
            return lin(x, residual=residual)
        return residual + linear_seq(x)

    def forward(self, hidden_states: torch.Tensor, context: torch.Tensor, time_embedding: torch.Tensor | None = None,
                cross_attention_kwargs: dict | None = None, *args, **kwargs) -> torch.Tensor:
        hidden_states = hidden_states + self.attn1(self.norm1(hidden_states))
        hidden_states = hidden_states + self.attn2(self.norm2(hidden_states), context=context, time_embedding=time_embedding,
                                                   **(cross_attention_kwargs or {}))
        hidden_states = hidden_states + self.ff(self.norm3(hidden_states))
        return hidden_states
