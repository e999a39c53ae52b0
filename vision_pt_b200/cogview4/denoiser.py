"""CogView4 (DiT) transformer block on the sm_100a kernels: the adaLN shift / scale / gate path that BASELINE.json's
north_star names.

Mirror of /root/reference/src/models/cogview4/denoiser.py: AdaLayerNormZero (148-200: one Linear -> 12 modulation chunks
in the order shift_msa, c_shift_msa, scale_msa, c_scale_msa, gate_msa, c_gate_msa, shift_mlp, c_shift_mlp, scale_mlp,
c_scale_mlp, gate_mlp, c_gate_mlp; separate affine-free LayerNorms for the image and the text stream), apply_rotary_emb
(203-218), SelfAttention over the joint [text | image] sequence (221-309), FeedForward (312-343), TransformerBlock
(346-423) and FinalAdaLayerNorm (486-523).  Names are the reference's (`norm1.linear`, `attn1.to_q`, `attn1.to_out.0`,
`ff.net.0.proj`, `ff.net.2`).  LayerNorm + modulate and gated residual are one fused pass each (`ln_modulate`,
`gate_residual`); QK LayerNorm, the half-split rotary embedding of the image tokens and gelu(tanh) are kernels of
csrc/blocks_ext.cuh; every Linear goes through the quant / PEFT registries.  The model around the block (patch embed,
text projection, RoPE table builder) is out of scope.
"""
from __future__ import annotations

import warnings
from typing import NamedTuple

import torch
import torch.nn as nn

from .. import ops
from ..modules.attention import AttentionImplementation
from ..modules.norm import FP32LayerNorm


class AdaLayerNormZeroOutput(NamedTuple):
    hidden_states: torch.Tensor
    img_gate_msa: torch.Tensor
    img_shift_mlp: torch.Tensor
    img_scale_mlp: torch.Tensor
    img_gate_mlp: torch.Tensor
    encoder_hidden_states: torch.Tensor
    cond_gate_msa: torch.Tensor
    cond_shift_mlp: torch.Tensor
    cond_scale_mlp: torch.Tensor
    cond_gate_mlp: torch.Tensor


class AdaLayerNormZero(nn.Module):
    def __init__(self, embedding_dim: int, dim: int) -> None:
        super().__init__()
        self.norm = FP32LayerNorm(dim, elementwise_affine=False, eps=1e-5)
        self.norm_context = FP32LayerNorm(dim, elementwise_affine=False, eps=1e-5)
        self.linear = nn.Linear(embedding_dim, 12 * dim, bias=True)

    def forward(self, hidden_states, encoder_hidden_states, time_embed) -> AdaLayerNormZeroOutput:
        emb = self.linear(time_embed)
        (shift_msa, c_shift_msa, scale_msa, c_scale_msa, gate_msa, c_gate_msa, shift_mlp, c_shift_mlp, scale_mlp, c_scale_mlp,
         gate_mlp, c_gate_mlp) = emb.chunk(12, dim=1)
        hidden_states = ops.ln_modulate(hidden_states, scale_msa, shift_msa, self.norm.eps)
        encoder_hidden_states = ops.ln_modulate(encoder_hidden_states, c_scale_msa, c_shift_msa, self.norm_context.eps)
        return AdaLayerNormZeroOutput(hidden_states, gate_msa, shift_mlp, scale_mlp, gate_mlp, encoder_hidden_states,
                                      c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp)


def apply_rotary_emb(inputs: torch.Tensor, freqs_cis: tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    """Reference signature ([B, H, S, hd] input, (cos, sin) of [S, hd]); runs the fused kernel on a token-major copy."""
    cos, sin = freqs_cis
    x4 = inputs.permute(0, 2, 1, 3)
    return ops.rope_half(x4, cos, sin, 0).permute(0, 2, 1, 3)


class SelfAttention(nn.Module):
    def __init__(self, hidden_dim: int, num_heads: int, bias: bool = True, attention_backend: AttentionImplementation = "eager"):
        super().__init__()
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.head_dim = hidden_dim // num_heads
        self.attention_backend = attention_backend
        self.to_q = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self.to_k = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self.to_v = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self.norm_q = FP32LayerNorm(self.head_dim, elementwise_affine=False, eps=1e-5)
        self.norm_k = FP32LayerNorm(self.head_dim, elementwise_affine=False, eps=1e-5)
        self.to_out = nn.ModuleList([nn.Linear(hidden_dim, hidden_dim, bias=bias)])

    def forward(self, hidden_states, encoder_hidden_states, image_rotary_emb=None):
        B, text_len, _ = encoder_hidden_states.shape
        x = torch.cat([encoder_hidden_states, hidden_states], dim=1)
        L = x.shape[1]
        H, hd = self.num_heads, self.head_dim
        # token-major [B, L, H, hd] throughout: the LayerNorm over head_dim and the rotary embedding are per (token, head)
        q = self.norm_q(self.to_q(x).view(B, L, H, hd))
        k = self.norm_k(self.to_k(x).view(B, L, H, hd))
        v = self.to_v(x).view(B, L, H, hd)
        if image_rotary_emb is not None:
            cos, sin = image_rotary_emb
            q = ops.rope_half(q, cos, sin, text_len)          # image tokens only (reference :271-282)
            k = ops.rope_half(k, cos, sin, text_len)
        else:
            warnings.warn("RoPE embeddings are not provided. ")
        o = ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), None, hd ** -0.5)
        o = self.to_out[0](o.permute(0, 2, 1, 3).reshape(B, L, H * hd))
        return o[:, text_len:], o[:, :text_len]


class _Activation(nn.Module):
    def __init__(self, kind: str):
        super().__init__()
        if kind not in ops.ACT_KINDS:
            raise NotImplementedError(f"activation {kind} has no fused kernel (silu, gelu, gelu_tanh / gelu_pytorch_tanh do)")
        self.kind = kind

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.activation(x, self.kind)


class FeedForward(nn.Module):
    def __init__(self, hidden_dim: int, mlp_scale: float = 4.0, activation_fn: str = "gelu_pytorch_tanh", bias: bool = True):
        super().__init__()
        self.inner_dim = int(hidden_dim * mlp_scale)
        self.net = nn.ModuleList([nn.ModuleDict({"proj": nn.Linear(hidden_dim, self.inner_dim, bias=bias)}),
                                  _Activation(activation_fn), nn.Linear(self.inner_dim, hidden_dim, bias=bias)])

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        return self.net[2](self.net[1](self.net[0]["proj"](hidden_states)))


class TransformerBlock(nn.Module):
    def __init__(self, hidden_dim: int = 2560, num_attention_heads: int = 64, time_embed_dim: int = 512,
                 attention_backend: AttentionImplementation = "eager"):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.norm1 = AdaLayerNormZero(time_embed_dim, hidden_dim)
        self.attn1 = SelfAttention(hidden_dim=hidden_dim, num_heads=num_attention_heads, bias=True,
                                   attention_backend=attention_backend)
        self.norm2 = FP32LayerNorm(hidden_dim, elementwise_affine=False, eps=1e-5)
        self.norm2_context = FP32LayerNorm(hidden_dim, elementwise_affine=False, eps=1e-5)
        self.ff = FeedForward(hidden_dim=hidden_dim)

    def forward(self, hidden_states, encoder_hidden_states, time_embed=None, image_rotary_emb=None):
        m = self.norm1(hidden_states, encoder_hidden_states, time_embed)
        attn_img, attn_txt = self.attn1(hidden_states=m.hidden_states, encoder_hidden_states=m.encoder_hidden_states,
                                        image_rotary_emb=image_rotary_emb)
        hidden_states = ops.gate_residual(hidden_states, attn_img, m.img_gate_msa)
        encoder_hidden_states = ops.gate_residual(encoder_hidden_states, attn_txt, m.cond_gate_msa)
        norm_img = ops.ln_modulate(hidden_states, m.img_scale_mlp, m.img_shift_mlp, self.norm2.eps)
        norm_txt = ops.ln_modulate(encoder_hidden_states, m.cond_scale_mlp, m.cond_shift_mlp, self.norm2_context.eps)
        hidden_states = ops.gate_residual(hidden_states, self.ff(norm_img), m.img_gate_mlp)
        encoder_hidden_states = ops.gate_residual(encoder_hidden_states, self.ff(norm_txt), m.cond_gate_mlp)
        return hidden_states, encoder_hidden_states


class FinalAdaLayerNorm(nn.Module):
    def __init__(self, hidden_dim: int, condition_dim: int, elementwise_affine: bool = False, eps: float = 1e-5,
                 bias: bool = True, hidden_act: str = "silu"):
        super().__init__()
        if elementwise_affine:
            raise NotImplementedError("FinalAdaLayerNorm is built for the affine-free norm CogView4 uses")
        self.linear = nn.Linear(condition_dim, 2 * hidden_dim, bias=bias)
        self.norm = FP32LayerNorm(hidden_dim, elementwise_affine=False, eps=eps)
        self.act = nn.SiLU() if hidden_act == "silu" else _Activation(hidden_act)

    def forward(self, hidden_states: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        condition = self.act(condition).to(hidden_states.dtype)        # [B, condition_dim]: per-sample glue
        scale, shift = self.linear(condition).chunk(2, dim=-1)
        return ops.ln_modulate(hidden_states, scale, shift, self.norm.eps)
