"""CPU oracle for the block families either side of the JiT block.  TEST INFRASTRUCTURE ONLY (see oracle/nf4.py for the
import rule).

Plain-PyTorch functions over a flat ``{reference state-dict name: tensor}`` dict, restating
  SDXL SelfAttention / CrossAttention / GeGLU / FeedForward / TransformerBlock   /root/reference/src/models/sdxl/denoiser.py:32-280
  CogView4 AdaLayerNormZero / apply_rotary_emb / SelfAttention / FeedForward / TransformerBlock / FinalAdaLayerNorm
                                                                                /root/reference/src/models/cogview4/denoiser.py:148-423, 486-523
  apply_pope                                                                    /root/reference/src/models/jit/extension/pope.py:6-38
  TREAD keep / route / re-insert                                                /root/reference/train/jit/class_to_image_tread.py:73-118
  U-JiT skip merge (Linear(2D -> D) over cat([x, skip]))                        /root/reference/src/models/jit/extension/uvit.py:30-147
Pinned against the reference's own modules imported live in the build container: tests/golden/make_golden_blocks.py wrote
tests/golden/block_family_vectors.pt from /root/reference, tests/test_oracle_golden.py replays them through this file.
Arithmetic follows the dtype of the inputs (bf16 in = the reference's rounding points, fp32 in = the error budget's zero).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .jit import attention, lora_linear


def _lin(P: dict, name: str, x, alpha: float):
    if f"{name}.lora_down.weight" in P:
        return lora_linear(x, P[f"{name}.linear.weight"], P.get(f"{name}.linear.bias"), P[f"{name}.lora_down.weight"],
                           P[f"{name}.lora_up.weight"], alpha)
    return lora_linear(x, P[f"{name}.weight"], P.get(f"{name}.bias"))


def _heads(t, H):
    B, L, D = t.shape
    return t.reshape(B, L, H, D // H).permute(0, 2, 1, 3)


def _merge(t):
    B, H, L, hd = t.shape
    return t.permute(0, 2, 1, 3).reshape(B, L, H * hd)


# ------------------------------------------------------------------------------------------------ SDXL
def sdxl_attention(P, prefix, x, ctx, num_heads, alpha):
    q = _lin(P, f"{prefix}.to_q", x, alpha)
    k = _lin(P, f"{prefix}.to_k", ctx, alpha)
    v = _lin(P, f"{prefix}.to_v", ctx, alpha)
    o = attention(_heads(q, num_heads), _heads(k, num_heads), _heads(v, num_heads)).to(x.dtype)
    return _lin(P, f"{prefix}.to_out.0", _merge(o), alpha)


def sdxl_block(P: dict, x, context, num_heads: int, alpha: float = 1.0, prefix: str = ""):
    ln = lambda n, t: F.layer_norm(t, (t.shape[-1],), P[f"{prefix}{n}.weight"].to(t.dtype), P[f"{prefix}{n}.bias"].to(t.dtype), 1e-5)
    h = ln("norm1", x)
    x = x + sdxl_attention(P, f"{prefix}attn1", h, h, num_heads, alpha)
    x = x + sdxl_attention(P, f"{prefix}attn2", ln("norm2", x), context, num_heads, alpha)
    h, gate = _lin(P, f"{prefix}ff.net.0.proj", ln("norm3", x), alpha).chunk(2, dim=-1)
    return x + _lin(P, f"{prefix}ff.net.2", h * F.gelu(gate), alpha)


# ------------------------------------------------------------------------------------------------ CogView4
def _ln32(x, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), None, None, eps).to(x.dtype)


def rotary_half(x, cos, sin):
    """apply_rotary_emb: x [B, H, S, hd]; cos / sin [S, hd]."""
    real, imag = x.reshape(*x.shape[:-1], 2, -1).unbind(-2)
    rot = torch.cat([-imag, real], dim=-1)
    return (x.float() * cos[None, None].to(x.device) + rot.float() * sin[None, None].to(x.device)).to(x.dtype)


def cogview4_block(P: dict, x, enc, time_embed, rotary, num_heads: int, alpha: float = 1.0, prefix: str = ""):
    emb = F.linear(time_embed, P[f"{prefix}norm1.linear.weight"].to(x.dtype), P[f"{prefix}norm1.linear.bias"].to(x.dtype))
    (shift_msa, c_shift_msa, scale_msa, c_scale_msa, gate_msa, c_gate_msa, shift_mlp, c_shift_mlp, scale_mlp, c_scale_mlp,
     gate_mlp, c_gate_mlp) = emb.chunk(12, dim=1)
    nx = _ln32(x) * (1 + scale_msa.unsqueeze(1)) + shift_msa.unsqueeze(1)
    ne = _ln32(enc) * (1 + c_scale_msa.unsqueeze(1)) + c_shift_msa.unsqueeze(1)
    T = enc.shape[1]
    h = torch.cat([ne, nx], dim=1)
    q = _heads(_lin(P, f"{prefix}attn1.to_q", h, alpha), num_heads)
    k = _heads(_lin(P, f"{prefix}attn1.to_k", h, alpha), num_heads)
    v = _heads(_lin(P, f"{prefix}attn1.to_v", h, alpha), num_heads)
    q, k = _ln32(q), _ln32(k)
    if rotary is not None:
        cos, sin = rotary
        q = torch.cat([q[:, :, :T], rotary_half(q[:, :, T:], cos, sin)], dim=2)
        k = torch.cat([k[:, :, :T], rotary_half(k[:, :, T:], cos, sin)], dim=2)
    o = _merge(attention(q, k, v).to(x.dtype))
    o = _lin(P, f"{prefix}attn1.to_out.0", o, alpha)
    x = x + o[:, T:] * gate_msa.unsqueeze(1)
    enc = enc + o[:, :T] * c_gate_msa.unsqueeze(1)
    nx = _ln32(x) * (1 + scale_mlp.unsqueeze(1)) + shift_mlp.unsqueeze(1)
    ne = _ln32(enc) * (1 + c_scale_mlp.unsqueeze(1)) + c_shift_mlp.unsqueeze(1)

    def ff(t):
        return _lin(P, f"{prefix}ff.net.2", F.gelu(_lin(P, f"{prefix}ff.net.0.proj", t, alpha), approximate="tanh"), alpha)

    x = x + ff(nx) * gate_mlp.unsqueeze(1)
    enc = enc + ff(ne) * c_gate_mlp.unsqueeze(1)
    return x, enc


def cogview4_final_norm(P: dict, x, cond, prefix: str = ""):
    c = F.silu(cond).to(x.dtype)
    scale, shift = F.linear(c, P[f"{prefix}linear.weight"].to(x.dtype), P[f"{prefix}linear.bias"].to(x.dtype)).chunk(2, dim=-1)
    return _ln32(x) * (1 + scale)[:, None, :] + shift[:, None, :]


# ------------------------------------------------------------------------------------------------ JiT extensions
def pope(x, freqs_cis, bias=None):
    """apply_pope: x [B, H, L, d] -> [B, H, L, 2d] (interleaved re / im), freqs_cis complex [L, d], bias [H, d] or None."""
    mag = F.softplus(x.float()).to(torch.complex64)
    fc = freqs_cis.view(1, 1, *freqs_cis.shape)
    if bias is not None:
        fc = fc * torch.polar(torch.ones_like(bias).float(), bias.float()).to(torch.complex64).view(1, bias.shape[0], 1, -1)
    return torch.view_as_real(mag * fc).flatten(3).type_as(x)


def tread_split(tokens, perm, num_keep: int):
    """keep_and_route_tokens: (keep, route) = tokens[:, perm[:num_keep]], tokens[:, perm[num_keep:]]."""
    return tokens[:, perm[:num_keep]], tokens[:, perm[num_keep:]]


def tread_merge(keep, route, perm):
    """Re-insertion at the end of routing: cat along tokens, undo the permutation (argsort(perm))."""
    return torch.cat([keep, route], dim=1)[:, torch.argsort(perm)]


def skip_merge(P: dict, name: str, x, skip, alpha: float = 1.0):
    """U-JiT long skip connection: Linear(2D -> D) over cat([x, skip], dim=-1)."""
    return _lin(P, name, torch.cat([x, skip], dim=-1), alpha)
