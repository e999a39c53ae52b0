"""CPU oracle for the JiT / DiT block path.  TEST INFRASTRUCTURE ONLY (see oracle/nf4.py for the import rule).

A plain-PyTorch restatement, written as functions over a flat ``{reference state-dict name: tensor}`` dict, of
  LoRALinear.forward                  /root/reference/src/modules/peft/lora.py:92-104
  FP32RMSNorm / FP32LayerNorm          /root/reference/src/modules/norm.py:9-27
  SingleAdaLayerNormZero / adaLN gate  /root/reference/src/modules/norm.py:70-90, src/models/cogview4/denoiser.py:182-187,401-420
  apply_rope / RopeEmbedder            /root/reference/src/models/jit/denoiser.py:98-287
  Attention / SwiGLU / JiTBlock / JiT  /root/reference/src/models/jit/denoiser.py:290-397, 480-543, 582-649, 969-1124
  scaled_dot_product_attention         /root/reference/src/modules/attention.py:98-129
  patchify / unpatchify                /root/reference/src/modules/patch.py:17-115, jit/denoiser.py:828-860
  rectified-flow loss of train_step    /root/reference/train/jit/class_to_image.py:106-242
Pinned against the reference's own modules imported live in the build container: tests/golden/make_golden.py wrote
tests/golden/*.pt from /root/reference, and tests/test_oracle_golden.py replays them through this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .nf4 import Nf4State, dequantize_nf4


# ------------------------------------------------------------------------------------------------ primitives
def dense_weight(w) -> torch.Tensor:
    return dequantize_nf4(w) if isinstance(w, Nf4State) else w


def lora_linear(x, w, bias, down=None, up=None, alpha: float = 1.0):
    """base(x) + lora_up(lora_down(x)) * (alpha / rank), each intermediate in x.dtype like the reference."""
    wd = dense_weight(w).to(x.dtype)
    out = F.linear(x, wd, None if bias is None else bias.to(x.dtype))
    if down is None:
        return out
    rank = down.shape[0]
    t = F.linear(x, down.to(x.dtype))
    u = F.linear(t, up.to(x.dtype))
    scale = torch.tensor(alpha, dtype=down.dtype) / rank
    return out + u * scale.to(x.dtype)


def rms_norm_fp32(x, weight, eps: float = 1e-6):
    return F.rms_norm(x.to(torch.float32), (x.shape[-1],), weight=weight, eps=eps).to(x.dtype)


def layer_norm_fp32(x, eps: float = 1e-5):
    return F.layer_norm(x.to(torch.float32), (x.shape[-1],), None, None, eps).to(x.dtype)


def adaln_modulate(x, scale, shift, eps: float = 1e-5):
    return layer_norm_fp32(x, eps) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def gate_residual(x, h, gate):
    return x + h * gate.unsqueeze(1)


def swiglu_gate(g, u):
    return F.silu(g) * u


# True: fp32 inputs stay fp32 in attention (a "ground truth" arm for error budgets); the reference itself casts fp32
# q/k/v to bf16 (src/modules/attention.py:113-118), which is the default here
ATTENTION_FP32 = False


def attention(q, k, v, key_mask=None):
    """q,k,v [B,H,L,hd]; key_mask [B,Lk] (1 = attend).  fp32 inputs are computed in bf16 like the reference."""
    if q.dtype == torch.float32 and ATTENTION_FP32:
        key_len = None if key_mask is None else key_mask.bool().sum(dim=1)
        return attention_explicit(q, k, v, key_len)
    if q.dtype == torch.float32:
        q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    mask = None
    if key_mask is not None:
        B, H, Lq, _ = q.shape
        mask = key_mask.bool().view(B, 1, 1, -1).expand(-1, H, Lq, -1)
    return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)


def attention_explicit(q, k, v, key_len=None, scale=None):
    """fp32 softmax(q k^T) v with per-sample valid key counts -- the kernel-independent form used for gradients."""
    B, H, Lq, hd = q.shape
    scale = hd ** -0.5 if scale is None else scale
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if key_len is not None:
        idx = torch.arange(k.shape[2], device=s.device).view(1, 1, 1, -1)
        s = s.masked_fill(idx >= key_len.view(B, 1, 1, 1), float("-inf"))
    return torch.matmul(torch.softmax(s, dim=-1), v.float())


def patchify(image, p: int, order: int = 0):
    B, C, H, W = image.shape
    x = image.view(B, C, H // p, p, W // p, p)
    x = x.permute(0, 2, 4, 1, 3, 5) if order == 0 else x.permute(0, 2, 4, 3, 5, 1)
    return x.reshape(B, (H // p) * (W // p), C * p * p)


def unpatchify(patches, C: int, H: int, W: int, p: int, order: int = 0):
    B = patches.shape[0]
    if order == 0:
        x = patches.reshape(B, H // p, W // p, C, p, p).permute(0, 3, 1, 4, 2, 5)
    else:
        x = patches.reshape(B, H // p, W // p, p, p, C).permute(0, 5, 1, 3, 2, 4)
    return x.reshape(B, C, H, W)


# ------------------------------------------------------------------------------------------------ RoPE
def rope_freqs_cis(cfg: dict, height: int, width: int, context_len: int, n_size: int = 6) -> torch.Tensor:
    """complex64 [L, head_dim/2], token order patches -> size -> time -> context."""
    p = cfg["patch_size"]
    hp, wp = height // p, width // p
    theta = cfg["rope_theta"]

    def axis_table(dim, positions):
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        ang = torch.outer(positions.to(torch.float64), inv).float()
        return torch.polar(torch.ones_like(ang), ang).to(torch.complex64)

    def tokens(gidx, a1, a2):
        cols = []
        for dim, pos in zip(cfg["rope_axes_dims"], (torch.full_like(a1, gidx), a1, a2)):
            cols.append(axis_table(dim, pos))
        return torch.cat(cols, dim=-1)

    ys = torch.arange(hp // 2 - hp, hp // 2).unsqueeze(1).repeat(1, wp).reshape(-1).float()
    xs = torch.arange(wp // 2 - wp, wp // 2).unsqueeze(0).repeat(hp, 1).reshape(-1).float()
    parts = [tokens(3.0, ys, xs)]
    for gidx, n in ((2.0, n_size), (1.0, cfg["num_time_tokens"]), (0.0, context_len)):
        j = torch.arange(n).float()
        parts.append(tokens(gidx, j, j))
    return torch.cat(parts, dim=0)


def apply_rope(x, freqs_cis):
    """x [B,H,L,hd], freqs_cis [L, hd/2] complex: rotate interleaved pairs in fp32, cast back."""
    B, H, L, hd = x.shape
    xc = torch.view_as_complex(x.float().reshape(B, H, L, hd // 2, 2))
    return torch.view_as_real(xc * freqs_cis.view(1, 1, L, hd // 2)).flatten(3).type_as(x)


# ------------------------------------------------------------------------------------------------ block / model
def _lin(P: dict, name: str, x, alpha: float):
    """Linear `name` of the parameter dict, with LoRA when `<name>.lora_down.weight` is present (peft key layout)."""
    if f"{name}.lora_down.weight" in P:
        return lora_linear(x, P[f"{name}.linear.weight"], P.get(f"{name}.linear.bias"), P[f"{name}.lora_down.weight"],
                           P[f"{name}.lora_up.weight"], alpha)
    return lora_linear(x, P[f"{name}.weight"], P.get(f"{name}.bias"))


def jit_block(P: dict, prefix: str, x, freqs_cis, key_mask, num_heads: int, alpha: float = 1.0, eps: float = 1e-6):
    B, L, D = x.shape
    hd = D // num_heads
    h = rms_norm_fp32(x, P[f"{prefix}norm1.weight"], eps)
    split = lambda t: t.view(B, L, num_heads, hd).permute(0, 2, 1, 3)
    q = split(_lin(P, f"{prefix}attn.to_q", h, alpha))
    k = split(_lin(P, f"{prefix}attn.to_k", h, alpha))
    v = split(_lin(P, f"{prefix}attn.to_v", h, alpha))
    q = apply_rope(rms_norm_fp32(q, P[f"{prefix}attn.q_norm.weight"], eps), freqs_cis)
    k = apply_rope(rms_norm_fp32(k, P[f"{prefix}attn.k_norm.weight"], eps), freqs_cis)
    o = attention(q, k, v, key_mask).to(x.dtype)
    o = o.permute(0, 2, 1, 3).contiguous().view(B, L, D)
    x = x + _lin(P, f"{prefix}attn.to_o", o, alpha)
    h = rms_norm_fp32(x, P[f"{prefix}norm2.weight"], eps)
    a = swiglu_gate(_lin(P, f"{prefix}mlp.w_1", h, alpha), _lin(P, f"{prefix}mlp.w_2", h, alpha))
    return x + _lin(P, f"{prefix}mlp.w_3", a, alpha)


def timestep_embedding(t, dim: int = 256):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    ang = t[:, None].float() * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)   # flip_sin_to_cos=True


def _embedder(P, prefix, t, dtype):
    h = F.linear(timestep_embedding(t).to(dtype), P[f"{prefix}mlp.0.weight"], P[f"{prefix}mlp.0.bias"])
    return F.linear(F.silu(h), P[f"{prefix}mlp.2.weight"], P[f"{prefix}mlp.2.bias"])


def jit_forward(P: dict, cfg: dict, image, timestep, context, original_size, target_size, crop_coords,
                context_mask=None, alpha: float = 1.0):
    """JiT.forward: in-context conditioning tokens, context re-appended at every block >= context_start_block."""
    B, _, height, width = image.shape
    D, heads, p = cfg["hidden_size"], cfg["num_heads"], cfg["patch_size"]
    dtype = P["context_embedder.weight"].dtype
    time_tokens = _embedder(P, "time_embedder.", timestep * cfg.get("timestep_scale", 1.0), dtype).unsqueeze(1) \
        + P["time_position_embeds"].unsqueeze(0)
    ctx = F.linear(context, P["context_embedder.weight"], P["context_embedder.bias"])
    sizes = torch.cat([original_size, target_size, crop_coords], dim=1).view(-1)
    size_tokens = _embedder(P, "image_size_embedder.", sizes, dtype).view(B, 6, D)
    patches = F.conv2d(image, P["patch_embedder.proj_1.weight"], None, stride=p)
    patches = F.conv2d(patches, P["patch_embedder.proj_2.weight"], P["patch_embedder.proj_2.bias"]).flatten(2).transpose(1, 2)
    n_patch, n_ctx = patches.shape[1], ctx.shape[1]
    dev = image.device                                   # the checker runs wherever its inputs live (CPU in tests/, GPU for the 200-step curve)
    freqs = rope_freqs_cis(cfg, height, width, n_ctx).to(dev)
    ones = torch.ones(B, n_patch + 6 + time_tokens.shape[1], device=dev)
    mask = torch.cat([ones, context_mask.float().to(dev) if context_mask is not None else torch.ones(B, n_ctx, device=dev)], dim=1)
    tokens = torch.cat([patches, size_tokens, time_tokens], dim=1)
    csb, fuse = cfg.get("context_start_block", 0), cfg.get("do_context_fuse", False)
    for i in range(cfg["depth"]):
        if i == csb or (not fuse and i >= csb):
            tokens = torch.cat([tokens, ctx], dim=1)
        L = tokens.shape[1]
        tokens = jit_block(P, f"blocks.{i}.", tokens, freqs[:L], mask[:, :L], heads, alpha)
        if not fuse and i >= csb:
            tokens = tokens[:, :-n_ctx, :]
    x = rms_norm_fp32(tokens[:, :n_patch], P["final_layer.norm_final.weight"], 1e-6)
    x = swiglu_gate(_lin(P, "final_layer.mlp.w_1", x, alpha), _lin(P, "final_layer.mlp.w_2", x, alpha))
    x = _lin(P, "final_layer.mlp.w_3", x, alpha)
    x = F.linear(x, P["final_layer.linear.weight"], P["final_layer.linear.bias"])
    return unpatchify(x, cfg.get("out_channels", 3), height, width, p, order=1)


def velocity_loss(pred_image, clean, noisy, timestep, clamp_eps: float = 0.05):
    """x-prediction scored as velocity MSE (treat_loss with model_pred='image', loss_target='velocity')."""
    denom = (1 - timestep.view(-1, 1, 1, 1)).clamp_min(clamp_eps)
    return F.mse_loss((pred_image - noisy) / denom, (clean - noisy) / denom, reduction="mean")
