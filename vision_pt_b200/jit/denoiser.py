"""JiT network on the sm_100a kernels.

Mirror of /root/reference/src/models/jit/denoiser.py: BottleneckPatchEmbed (17-67), TimestepEmbedder (70-95),
RopeEmbedder table math (114-287), Attention (290-397), SwiGLU (480-506), FinalLayer (509-543), JiTBlock (582-649),
JiT (652-1124).  Module and parameter names are the reference's, so its checkpoints and PEFT key patterns apply.
Differences are in HOW, not WHAT: the RoPE table is built once per (H, W) bucket and kept on the device; the bool
key-padding mask becomes per-sample key lengths; a block runs as a fixed sequence of fused kernels with a hand-written
backward (`JiTBlockFn`) whenever only LoRA parameters train.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..modules.attention import prefix_key_lengths
from ..modules.norm import get_norm_layer
from ..modules.peft import LoRALinear
from ..modules.quant import NF4Linear
from .config import DenoiserConfig


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, flip_sin_to_cos: bool = False,
                           downscale_freq_shift: float = 1, scale: float = 1, max_period: int = 10000) -> torch.Tensor:
    """Sinusoidal embedding, /root/reference/src/modules/timestep/embedding.py:10-62 (per-step [B]-sized glue)."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = scale * (timesteps[:, None].float() * torch.exp(exponent)[None, :])
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def _kernel_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor | None:
    """x W^T + b through the tcgen05 kernel for a FROZEN bf16 weight (patch embed, final layer, context embedder);
    None when the layer does not qualify (trainable, other dtype, CPU) and the caller falls back to torch."""
    if (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and not weight.requires_grad
            and (bias is None or not bias.requires_grad)):
        # a ragged in_features (the final layer's 3413- / 2730-wide SwiGLU) goes through a row-padded copy of the weight:
        # torch's matmul falls to an unaligned legacy kernel there (1.2 ms per call at JiT-H sizes)
        return ops.nf4_lora_linear(x, ops.padded_weight(weight), bias)
    return None


class BottleneckPatchEmbed(nn.Module):
    def __init__(self, patch_size=16, in_channels=3, bottleneck_dim=128, hidden_dim=768, bias=True):
        super().__init__()
        self.patch_size = patch_size
        self.proj_1 = nn.Conv2d(in_channels, bottleneck_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.proj_2 = nn.Conv2d(bottleneck_dim, hidden_dim, kernel_size=1, stride=1, bias=bias)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        w1, w2 = self.proj_1.weight, self.proj_2.weight
        if image.is_cuda and image.dtype == torch.bfloat16 and not image.requires_grad:
            # a stride-p conv over non-overlapping patches = patchify in (c, py, px) order + GEMM (reference :51-67)
            patches = ops.patchify_op(image, self.patch_size, order=0)
            h = _kernel_linear(patches, w1.view(w1.shape[0], -1), None)
            if h is not None:
                y = _kernel_linear(h, w2.view(w2.shape[0], -1), self.proj_2.bias)
                if y is not None:
                    return y
        return self.proj_2(self.proj_1(image)).flatten(2).transpose(1, 2)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_dim: int, freq_embedding_size: int = 256):
        super().__init__()
        self.freq_embedding_size = freq_embedding_size
        self.mlp = nn.Sequential(nn.Linear(freq_embedding_size, hidden_dim, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_dim, hidden_dim, bias=True))

    def forward(self, timestep: torch.Tensor) -> torch.Tensor:
        freq = get_timestep_embedding(timestep, self.freq_embedding_size, flip_sin_to_cos=True, downscale_freq_shift=0)
        return self.mlp(freq.to(dtype=self.mlp[0].weight.dtype))


def rope_table(cfg: DenoiserConfig, height: int, width: int, context_len: int, num_size_tokens: int = 6) -> torch.Tensor:
    """(cos, sin) of every token's rotation, [L, head_dim/2, 2] fp32, token order patches -> size -> time -> context.

    Same numbers as RopeEmbedder (reference denoiser.py:150-287): angle = fp32(pos * theta^(-2k/dim)) computed in
    float64, one group of frequencies per axis; patches sit at (3, y - h/2, x - w/2) zero-centred, the size / time /
    context tokens at (2, j, j), (1, j, j), (0, j, j)."""
    p = cfg.patch_size
    hp, wp = height // p, width // p
    ys = torch.arange(hp // 2 - hp, hp // 2, dtype=torch.float64).unsqueeze(1).expand(hp, wp).reshape(-1)
    xs = torch.arange(wp // 2 - wp, wp // 2, dtype=torch.float64).unsqueeze(0).expand(hp, wp).reshape(-1)
    pos = [torch.stack([torch.full_like(ys, 3.0), ys, xs], dim=1)]
    for gidx, n in ((2, num_size_tokens), (1, cfg.num_time_tokens), (0, context_len)):
        j = torch.arange(n, dtype=torch.float64)
        pos.append(torch.stack([torch.full_like(j, float(gidx)), j, j], dim=1))
    pos = torch.cat(pos, dim=0)                                  # [L, 3]
    angles = []
    for a, dim in enumerate(cfg.rope_axes_dims):
        freqs = 1.0 / (cfg.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.float64) / dim))
        angles.append(torch.outer(pos[:, a], freqs).float())
    ang = torch.cat(angles, dim=1)                               # [L, head_dim/2] fp32
    cis = torch.polar(torch.ones_like(ang), ang)                 # the reference builds the table with torch.polar
    return torch.stack([cis.real, cis.imag], dim=-1).contiguous()


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True, qk_norm=True, attn_dropout=0.0, proj_dropout=0.0, eps=1e-6,
                 norm_type="rms"):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.q_norm = get_norm_layer(norm_type, self.head_dim, eps=eps) if qk_norm else nn.Identity()
        self.k_norm = get_norm_layer(norm_type, self.head_dim, eps=eps) if qk_norm else nn.Identity()
        self.to_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_k = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_v = nn.Linear(dim, dim, bias=qkv_bias)
        self.attn_dropout = nn.Dropout(attn_dropout)
        self.to_o = nn.Linear(dim, dim)
        self.proj_dropout = nn.Dropout(proj_dropout)

    def forward(self, hidden_states, cos_sin, seqlens=None):
        """Per-op path (autograd composes the kernels): QKV -> fused QK-norm + RoPE -> attention -> to_o."""
        B, L, D = hidden_states.shape
        H, hd = self.num_heads, self.head_dim
        q = self.to_q(hidden_states).view(B, L, H, hd)
        k = self.to_k(hidden_states).view(B, L, H, hd)
        v = self.to_v(hidden_states).view(B, L, H, hd)
        q = ops.qknorm_rope(q, self.q_norm.weight, cos_sin, self.q_norm.eps)
        k = ops.qknorm_rope(k, self.k_norm.weight, cos_sin, self.k_norm.eps)
        o = ops.attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), seqlens, hd ** -0.5)
        o = o.permute(0, 2, 1, 3).reshape(B, L, D)
        return self.proj_dropout(self.to_o(o))


class SwiGLU(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0, bias=True):
        super().__init__()
        hidden_dim = int(hidden_dim * 2 / 3)
        self.w_1 = nn.Linear(dim, hidden_dim, bias=bias)
        self.w_2 = nn.Linear(dim, hidden_dim, bias=bias)
        self.w_3 = nn.Linear(hidden_dim, dim, bias=bias)
        self.ffn_dropout = nn.Dropout(dropout)

    @staticmethod
    def _lin(layer, x):
        if type(layer) is nn.Linear:
            y = _kernel_linear(x, layer.weight, layer.bias)
            if y is not None:
                return y
        return layer(x)

    def forward(self, hidden_states):
        if all(type(l) is nn.Linear for l in (self.w_1, self.w_2, self.w_3)) and self.ffn_dropout.p == 0.0:
            y = ops.dense_swiglu(hidden_states, self.w_1.weight, self.w_1.bias, self.w_2.weight, self.w_2.bias,
                                 self.w_3.weight, self.w_3.bias)
            if y is not None:
                return y
        a = ops.swiglu(self._lin(self.w_1, hidden_states), self._lin(self.w_2, hidden_states))
        return self._lin(self.w_3, self.ffn_dropout(a))


class FinalLayer(nn.Module):
    def __init__(self, hidden_dim, mlp_ratio, patch_size, out_channels, eps=1e-6, norm_type="rms"):
        super().__init__()
        self.norm_final = get_norm_layer(norm_type, hidden_dim, eps=eps)
        self.mlp = SwiGLU(dim=hidden_dim, hidden_dim=int(hidden_dim * mlp_ratio), dropout=0.0, bias=True)
        self.linear = nn.Linear(hidden_dim, patch_size * patch_size * out_channels, bias=True)

    def forward(self, hidden_states):
        return SwiGLU._lin(self.linear, self.mlp(self.norm_final(hidden_states)))


class BottleneckFinalLayer(nn.Module):
    def __init__(self, hidden_dim, bottleneck_dim, patch_size, out_channels, norm_type="rms"):
        super().__init__()
        self.norm_final = get_norm_layer(norm_type, hidden_dim, eps=1e-6)
        self.proj_1 = nn.Linear(hidden_dim, bottleneck_dim, bias=False)
        self.proj_2 = nn.Linear(bottleneck_dim, patch_size * patch_size * out_channels, bias=True)

    def forward(self, hidden_states):
        return self.proj_2(self.proj_1(self.norm_final(hidden_states)))


# ------------------------------------------------------------------------------------------------- fused block
# Called as hook(block_index) at the end of every fused block's backward, once the block's LoRA gradients sit in the flat
# buffer: the data-parallel trainer issues the all-reduce of finished gradient chunks from here (train.py).
BLOCK_BACKWARD_HOOK = None


class _Lin:
    """What one of the seven block linears contributes to the fused sequence."""
    __slots__ = ("w", "bias", "down", "up", "scale", "rank")

    def __init__(self, layer: nn.Module):
        lora = layer if isinstance(layer, LoRALinear) and layer.enabled else None
        base = layer.linear if isinstance(layer, LoRALinear) else layer
        self.w = base.quant_state if isinstance(base, NF4Linear) else base.weight
        self.bias = base.bias
        self.down = lora.lora_down.weight if lora is not None else None
        self.up = lora.lora_up.weight if lora is not None else None
        self.scale = lora.scale if lora is not None else 1.0
        self.rank = lora.rank if lora is not None else 0


def _linear_ok(layer: nn.Module) -> bool:
    base = layer.linear if isinstance(layer, LoRALinear) else layer
    if isinstance(layer, LoRALinear) and layer.enabled and not layer.fusable:
        return False
    if isinstance(base, NF4Linear):
        ok = base.is_quantized
    else:
        ok = (type(base) is nn.Linear and base.weight.dtype == torch.bfloat16 and not base.weight.requires_grad
              and base.in_features % 8 == 0)
    return ok and (base.bias is None or not base.bias.requires_grad)


class JiTBlockFn(torch.autograd.Function):
    """One JiTBlock (reference denoiser.py:633-649) as a fixed sequence of about 11 forward / 16 backward kernel launches
    (q | k | v one sectioned GEMM, SwiGLU in the w_2 / w_3-dX epilogues, the batched NF4 dequantisation prefetched on a side
    stream, LoRA parameter gradients of the whole block in one launch).

    Gradients: the residual stream and the LoRA matrices.  Everything else in the block is frozen on this path."""

    NAMES = ("to_q", "to_k", "to_v", "to_o", "w_1", "w_2", "w_3")

    @staticmethod
    def forward(ctx, x, cos_sin, seqlens, spec, *lora):
        lins, n1w, n2w, qnw, knw, H, eps, tail, n_keep, _, block_index, nxt, _, qkv_bias = spec
        B, L, D = x.shape
        M = B * L
        x2 = ops.packed_tokens(x).reshape(M, D)
        pads = [ops._pad_rank(l.down, l.up) for l in lins]

        # all seven NF4 weights of the block dequantised by one launch into L2-resident slots -- by the previous block's
        # prefetch when there was one; and the NEXT block's launch goes out now, on the side stream
        slots = None
        if M >= ops.NF4_SCRATCH_MIN_M and ops.NF4_GEMM_MODE != "prologue":
            slots = ops.PREFETCH.take(("f", block_index))
            if slots is None:
                slots = ops.dequant_block([l.w for l in lins], [p_[0] for p_ in pads], [p_[1] for p_ in pads], transposed=False,
                                          arena_tag=("pf", torch.cuda.current_stream().cuda_stream, block_index & 1))
            if nxt is not None and slots is not None:
                ops.PREFETCH.issue(("f", block_index + 1), [l.w for l in nxt], [None] * len(nxt), [None] * len(nxt), False,
                                   (block_index + 1) & 1)

        def lin(i, inp, residual=None, epilogue=0):
            l = lins[i]
            return ops.linear_raw(inp, l.w, l.bias, pads[i][0], pads[i][1], l.scale, residual, want_side=True,
                                  scratch=slots[i] if slots is not None else None, epilogue=epilogue)

        h1, rstd1 = ops.rmsnorm_fwd_raw(x2, n1w, eps)
        wqkv = None
        if slots is not None and qkv_bias is not False and ops.FUSE_QKV and D % 128 == 0 and all(
                pads[i][0] is not None and lins[i].scale == lins[0].scale and tuple(lins[i].w.shape) == (D, D) for i in range(3)):
            wqkv = ops.fused_rows_view(slots[0:3], D, D)          # the three forward slots ARE the stacked [3D, D] weight
        if wqkv is not None:
            # q | k | v as ONE GEMM over h1 (three sections, each with its own LoRA pair): h1 is read once, one launch
            qkv, t_qkv = ops.linear_raw(h1, wqkv, qkv_bias, ops.stacked([pads[i][0] for i in range(3)]),
                                        ops.stacked([pads[i][1] for i in range(3)]), lins[0].scale, None, want_side=True,
                                        n_sections=3, tape_slot=slots[0])
            q_pre, k_pre, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
            t_q, t_k, t_v = t_qkv[0:16], t_qkv[16:32], t_qkv[32:48]
        else:
            q_pre, t_q = lin(0, h1)
            k_pre, t_k = lin(1, h1)
            v, t_v = lin(2, h1)
        q = ops.qknorm_rope_fwd_raw(q_pre, qnw, cos_sin, H, L, eps)
        k = ops.qknorm_rope_fwd_raw(k_pre, knw, cos_sin, H, L, eps)
        hd = D // H
        as4 = lambda t: t.view(B, L, H, hd).permute(0, 2, 1, 3)
        o4, lse2 = ops.attn_fwd_raw(as4(q), as4(k), as4(v), seqlens, hd ** -0.5)
        o2 = o4.permute(0, 2, 1, 3).reshape(M, D)
        x1, t_o = lin(3, o2, x2)
        h2, rstd2 = ops.rmsnorm_fwd_raw(x1, n2w, eps)
        g, t_g = lin(4, h2)
        if slots is not None and ops.FUSE_SWIGLU:
            a, u, t_u = lin(5, h2, g, epilogue=1)        # silu(g) * u in the w_2 GEMM's epilogue
        else:
            u, t_u = lin(5, h2)
            a = ops.swiglu_fwd_raw(g, u)
        y, t_3 = lin(6, a, x1)

        ctx.spec, ctx.pads, ctx.dims = spec, pads, (B, L, D)
        ctx.seqlens, ctx.cos_sin = seqlens, cos_sin
        ctx.save_for_backward(x2, rstd1, h1, q_pre, k_pre, v, q, k, o2, lse2, x1, rstd2, h2, g, u, a,
                              t_q, t_k, t_v, t_o, t_g, t_u, t_3)
        y3 = y.view(B, L, D)
        if tail is not None:
            # JiT re-appends the ORIGINAL context tokens in front of every block >= context_start_block and strips the
            # block's outputs for them (reference denoiser.py:1092-1113): with the slots kept in the buffer that is one small
            # strided copy here (and zeroing their gradient in backward) instead of a torch.cat of all tokens per block
            ops.copy_token_slots(y3, n_keep, tail)
        return y3

    @staticmethod
    def backward(ctx, dy):
        (x2, rstd1, h1, q_pre, k_pre, v, q, k, o2, lse2, x1, rstd2, h2, g, u, a,
         t_q, t_k, t_v, t_o, t_g, t_u, t_3) = ctx.saved_tensors
        lins, n1w, n2w, qnw, knw, H, eps, tail, n_keep, fresh_in, block_index, _, prv, _ = ctx.spec
        pads = ctx.pads
        B, L, D = ctx.dims
        M = B * L
        dev = dy.device
        dy2 = ops.packed_tokens(dy).reshape(M, D)      # e.g. the [:, :pre_ctx] slice the context concatenation hands back
        grads: list = [None] * 14

        slots = None
        if M >= ops.NF4_SCRATCH_MIN_M and ops.NF4_GEMM_MODE != "prologue":
            slots = ops.PREFETCH.take(("b", block_index))
            if slots is None:
                slots = ops.dequant_block([l.w for l in lins], [p_[0] for p_ in pads], [p_[1] for p_ in pads], transposed=True,
                                          arena_tag=("pf", torch.cuda.current_stream().cuda_stream, block_index & 1))
            if prv is not None and slots is not None:
                ppads = [ops._pad_rank(l.down, l.up) for l in prv]
                ops.PREFETCH.issue(("b", block_index - 1), [l.w for l in prv], [p_[0] for p_ in ppads], [p_[1] for p_ in ppads], True,
                                   (block_index - 1) & 1)

        def back(i, dout, residual=None, epilogue=0, in2=None):
            l = lins[i]
            return ops.linear_raw(dout, l.w, None, pads[i][0], pads[i][1], l.scale, residual, want_side=True, backward=True,
                                  scratch=slots[i] if slots is not None else None, epilogue=epilogue, in2=in2)

        # LoRA parameter gradients: collected over the block and reduced by ONE batched launch at the end; linears that
        # share their input (q/k/v <- h1, w_1/w_2 <- h2) form one item so the activation is read once
        want = [lins[i].down is not None and (ctx.needs_input_grad[4 + 2 * i] or ctx.needs_input_grad[5 + 2 * i])
                for i in range(7)]
        direct = all((not want[i]) or (ops.grad_sink(lins[i].down) is not None and ops.grad_sink(lins[i].up) is not None
                                       and lins[i].rank == ops.RANK) for i in range(7))
        items: list = []
        shared: dict = {}

        def lora_grads(i, dout, t_side, inp, dt_side):
            l = lins[i]
            if not want[i]:
                return
            if not direct:
                grads[2 * i], grads[2 * i + 1] = ops.lora_param_grads(dout, t_side, inp, dt_side, l.down, l.up, l.rank)
                return
            items.append((dout, [t_side], [ops.grad_sink(l.up)], False))
            grp = shared.get(inp.data_ptr())
            if grp is None or len(grp[1]) == 3:
                grp = (inp, [], [], True)
                shared[inp.data_ptr()] = grp
                items.append(grp)
            grp[1].append(dt_side)
            grp[2].append(ops.grad_sink(l.down))

        # MLP branch
        if slots is not None and ops.FUSE_SWIGLU:
            dg, du, dt_3 = back(6, dy2, g, epilogue=2, in2=u)   # the SwiGLU backward in the w_3 dX GEMM's epilogue
        else:
            da, dt_3 = back(6, dy2)
            dg, du = ops.swiglu_bwd_raw(da, g, u)
        lora_grads(6, dy2, t_3, a, dt_3)
        dh2, dt_g = back(4, dg)
        dh2, dt_u = back(5, du, dh2)
        lora_grads(4, dg, t_g, h2, dt_g)
        lora_grads(5, du, t_u, h2, dt_u)
        dx1 = ops.rmsnorm_bwd_raw(dh2, x1, n2w, rstd2, dy2, eps)
        # attention branch
        do2, dt_o = back(3, dx1)
        lora_grads(3, dx1, t_o, o2, dt_o)
        hd = D // H
        as4 = lambda t: t.view(B, L, H, hd).permute(0, 2, 1, 3)
        dq4, dk4, dv4 = ops.attn_bwd_raw(as4(q), as4(k), as4(v), as4(o2), as4(do2), lse2, ctx.seqlens, hd ** -0.5)
        dq_post = dq4.permute(0, 2, 1, 3).reshape(M, D)          # fp32, token-major memory: views, no copies
        dk_post = dk4.permute(0, 2, 1, 3).reshape(M, D)
        dv2 = dv4.permute(0, 2, 1, 3).reshape(M, D)
        dq_pre = ops.qknorm_rope_bwd_raw(dq_post, q_pre, qnw, ctx.cos_sin, H, L, eps)
        dk_pre = ops.qknorm_rope_bwd_raw(dk_post, k_pre, knw, ctx.cos_sin, H, L, eps)
        dh1, dt_q = back(0, dq_pre)
        dh1, dt_k = back(1, dk_pre, dh1)
        dh1, dt_v = back(2, dv2, dh1)
        lora_grads(0, dq_pre, t_q, h1, dt_q)
        lora_grads(1, dk_pre, t_k, h1, dt_k)
        lora_grads(2, dv2, t_v, h1, dt_v)
        if items:
            ops.lora_grad_batch(items)
        if BLOCK_BACKWARD_HOOK is not None:
            BLOCK_BACKWARD_HOOK(block_index)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.rmsnorm_bwd_raw(dh1, x2, n1w, rstd1, dx1, eps).view(B, L, D)
            if fresh_in:
                # this block's input had its context slots REPLACED by the previous block's epilogue: no gradient flows
                # through them into that block (the reference strips those rows, denoiser.py:1111-1113)
                ops.copy_token_slots(dx, n_keep, None)
        return (dx, None, None, None, *grads)


class JiTBlock(nn.Module):
    def __init__(self, hidden_dim, num_heads, mlp_ratio=4.0, attn_dropout=0.0, proj_dropout=0.0, ffn_dropout=0.0,
                 qkv_bias=True, qk_norm=True, bias=True, eps=1e-6, positional_encoding="rope", norm_type="rms"):
        super().__init__()
        if positional_encoding != "rope":
            raise NotImplementedError("PoPE attention is an experimental variant outside the B200 hot path")
        self.eps = eps
        self.norm1 = get_norm_layer(norm_type, hidden_dim, eps=eps)
        self.attn = Attention(dim=hidden_dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_norm=qk_norm,
                              attn_dropout=attn_dropout, proj_dropout=proj_dropout, eps=eps, norm_type=norm_type)
        self.norm2 = get_norm_layer(norm_type, hidden_dim, eps=eps)
        self.mlp = SwiGLU(dim=hidden_dim, hidden_dim=int(hidden_dim * mlp_ratio), dropout=ffn_dropout, bias=bias)
        self.use_fused = True
        self.block_index = -1            # position in JiT.blocks (set by JiT): tells the backward hook how far backward is

    def _linears(self):
        return [self.attn.to_q, self.attn.to_k, self.attn.to_v, self.attn.to_o, self.mlp.w_1, self.mlp.w_2, self.mlp.w_3]

    def _stacked_qkv_bias(self, lins):
        """[3D] bias of the q | k | v call (frozen: built once per version of the three biases); None = no biases,
        False = mixed (the fused call is not taken)."""
        bs = [lins[i].bias for i in range(3)]
        if all(b is None for b in bs):
            return None
        if any(b is None for b in bs):
            return False
        key = tuple((b.data_ptr(), b._version, b.device) for b in bs)
        hit = self.__dict__.get("_qkv_bias_cache")
        if hit is None or hit[0] != key:
            hit = (key, torch.cat([b.detach().to(torch.bfloat16) for b in bs]))
            self.__dict__["_qkv_bias_cache"] = hit
        return hit[1]

    def fused_eligible(self, x: torch.Tensor) -> bool:
        norms = [self.norm1, self.norm2, self.attn.q_norm, self.attn.k_norm]
        return (self.use_fused and x.is_cuda and x.dtype == torch.bfloat16 and self.attn.head_dim in ops.QKNORM_HEAD_DIMS
                and all(isinstance(n, nn.RMSNorm) and n.weight is not None and not n.weight.requires_grad for n in norms)
                and all(_linear_ok(l) for l in self._linears())
                and self.attn.attn_dropout.p == 0 and self.attn.proj_dropout.p == 0 and self.mlp.ffn_dropout.p == 0)

    def forward(self, hidden_states, cos_sin, seqlens=None, ctx_tail=None, n_keep=0, fresh_in=False, prefetch=False):
        """ctx_tail [B, L - n_keep, D]: when given, rows n_keep.. of the OUTPUT are replaced by it (fresh context tokens for
        the next block).  fresh_in: rows n_keep.. of the INPUT were produced that way by the previous block, so the
        gradient with respect to them is dropped."""
        if self.fused_eligible(hidden_states):
            lins = [_Lin(l) for l in self._linears()]
            bf = lambda w: w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)
            # neighbours whose dequantisation this block may start early (ops.DequantPrefetcher); only all-NF4, LoRA rank <= 16
            # neighbours qualify, and only when blocks run strictly in order (no per-block recomputation)
            nxt = prv = None
            if ops.PREFETCH_DEQUANT and prefetch:
                nb, pb = self.__dict__.get("_next_block"), self.__dict__.get("_prev_block")
                ok = lambda b_: b_ is not None and b_.fused_eligible(hidden_states) and all(
                    isinstance(l.w, ops.Nf4Tensors) and l.rank <= ops.RANK for l in (_Lin(m) for m in b_._linears()))
                if ok(nb):
                    nxt = [_Lin(m) for m in nb._linears()]
                if ok(pb) and torch.is_grad_enabled():
                    prv = [_Lin(m) for m in pb._linears()]
            spec = (lins, bf(self.norm1.weight), bf(self.norm2.weight), bf(self.attn.q_norm.weight),
                    bf(self.attn.k_norm.weight), self.attn.num_heads, self.eps,
                    ctx_tail.detach() if ctx_tail is not None else None, n_keep, bool(fresh_in), self.block_index, nxt, prv,
                    self._stacked_qkv_bias(lins))
            lora = []
            for l in lins:
                lora += [l.down, l.up]
            return JiTBlockFn.apply(hidden_states, cos_sin, seqlens, spec, *lora)
        hidden_states = hidden_states + self.attn(self.norm1(hidden_states), cos_sin, seqlens)
        out = hidden_states + self.mlp(self.norm2(hidden_states))
        if ctx_tail is not None:
            out = torch.cat([out[:, :n_keep], ctx_tail.detach()], dim=1)     # autograd drops the replaced rows' gradient
        return out


class JiT(nn.Module):
    def __init__(self, config: DenoiserConfig):
        super().__init__()
        self.config = config
        assert (config.hidden_size // config.num_heads) == sum(config.rope_axes_dims), \
            "The sum of rope_axes_dims must equal to hidden_size / num_heads = head_dim."
        if config.positional_encoding != "rope":
            raise NotImplementedError("only the RoPE variant is on the B200 hot path")
        self.patch_embedder = BottleneckPatchEmbed(config.patch_size, config.in_channels, config.bottleneck_dim,
                                                   config.hidden_size, bias=True)
        self.time_embedder = TimestepEmbedder(config.hidden_size, 256)
        self.time_position_embeds = nn.Parameter(torch.randn(config.num_time_tokens, config.hidden_size))
        self.image_size_embedder = TimestepEmbedder(config.hidden_size, 256)
        self.context_embedder = nn.Linear(config.context_dim, config.hidden_size, bias=True)
        self.blocks = nn.ModuleList([
            JiTBlock(config.hidden_size, config.num_heads, config.mlp_ratio, config.attn_dropout, config.proj_dropout, 0.0,
                     True, True, True, 1e-6, config.positional_encoding, config.norm_type)
            for _ in range(config.depth)])
        for i, blk in enumerate(self.blocks):
            blk.block_index = i
            # plain references (not sub-modules): who runs before / after this block, for the dequantisation prefetch
            blk.__dict__["_prev_block"] = self.blocks[i - 1] if i > 0 else None
            blk.__dict__["_next_block"] = self.blocks[i + 1] if i + 1 < len(self.blocks) else None
        if config.use_output_bottleneck:
            self.final_layer = BottleneckFinalLayer(config.hidden_size, config.bottleneck_dim, config.patch_size,
                                                    config.in_channels, norm_type="rms")
        else:
            self.final_layer = FinalLayer(config.hidden_size, config.mlp_ratio, config.patch_size, config.in_channels,
                                          eps=1e-6, norm_type="rms")
        self.gradient_checkpointing = False
        self._rope_cache: dict[tuple, torch.Tensor] = {}

    def initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear) and not isinstance(m, NF4Linear):
                nn.init.normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.LayerNorm, nn.RMSNorm)) and m.weight is not None:
                nn.init.ones_(m.weight)
        for conv in (self.patch_embedder.proj_1, self.patch_embedder.proj_2):
            nn.init.normal_(conv.weight.view(conv.weight.shape[0], -1), std=0.02)
        if self.patch_embedder.proj_2.bias is not None:
            nn.init.zeros_(self.patch_embedder.proj_2.bias)
        nn.init.normal_(self.time_position_embeds, std=0.02)

    def set_gradient_checkpointing(self, enable: bool = True):
        self.gradient_checkpointing = enable

    def rope_cos_sin(self, height: int, width: int, context_len: int, device) -> torch.Tensor:
        key = (height, width, context_len, str(device))
        t = self._rope_cache.get(key)
        if t is None:
            t = rope_table(self.config, height, width, context_len).to(device)
            self._rope_cache[key] = t
        return t

    def _unpatchify(self, patches, height, width):
        B = patches.shape[0]
        if patches.is_cuda and patches.element_size() == 2:
            return ops.unpatchify_op(patches, self.config.out_channels, height, width, self.config.patch_size, order=1)
        p, c = self.config.patch_size, self.config.out_channels
        x = patches.view(B, height // p, width // p, p, p, c).permute(0, 5, 1, 3, 2, 4)
        return x.reshape(B, c, height, width)

    def unpatchify(self, patches, height, width):
        if self.config.use_pixel_shuffle:
            if patches.is_cuda and patches.element_size() == 2:
                return ops.unpatchify_op(patches, self.config.out_channels, height, width, self.config.patch_size, order=0)
            p = self.config.patch_size
            x = patches.view(patches.shape[0], height // p, width // p, -1).permute(0, 3, 1, 2)
            return F.pixel_shuffle(x, upscale_factor=p)
        return self._unpatchify(patches, height, width)

    def get_imagesize_embed(self, original_size, target_size, crop_coords):
        size_info = torch.cat([original_size, target_size, crop_coords], dim=1).view(-1)
        return self.image_size_embedder(size_info).view(-1, 6, self.config.hidden_size)

    def forward_block(self, block, tokens, cos_sin, seqlens, ctx_tail=None, n_keep=0, fresh_in=False):
        if self.gradient_checkpointing and self.training:
            import torch.utils.checkpoint as checkpoint
            return checkpoint.checkpoint(block, tokens, cos_sin, seqlens, ctx_tail, n_keep, fresh_in, use_reentrant=False)
        return block(tokens, cos_sin, seqlens, ctx_tail, n_keep, fresh_in, True)    # blocks run in order: prefetch allowed

    def forward(self, image, timestep, context, original_size, target_size, crop_coords, context_mask=None):
        cfg = self.config
        B, _, height, width = image.shape
        time_embed = self.time_embedder(timestep * cfg.timestep_scale)
        time_tokens = time_embed.unsqueeze(1) + self.time_position_embeds.unsqueeze(0)
        n_time = time_tokens.shape[1]
        context_embed = SwiGLU._lin(self.context_embedder, context)
        ctx_len = context_embed.shape[1]
        size_embed = self.get_imagesize_embed(original_size, target_size, crop_coords)
        n_size = size_embed.shape[1]
        patches = self.patch_embedder(image)
        n_patch = patches.shape[1]
        pre_ctx = n_patch + n_size + n_time

        cos_sin = self.rope_cos_sin(height, width, ctx_len, image.device)
        # key-padding mask -> per-sample key length (valid context tokens come first, reference class_encoder.py:72-81)
        # -- a PREFIX mask: anything else is refused by prefix_key_lengths, not silently treated as one)
        if context_mask is not None:
            # JiT's own mask convention is "1: attend, 0: ignore" in any dtype (reference denoiser.py:375-381: mask.bool())
            seq_ctx = (pre_ctx + prefix_key_lengths(context_mask.to(image.device) != 0)).contiguous()
        else:
            seq_ctx = None

        ops.PREFETCH.reset()                         # nothing left over from an abandoned pass
        tokens = torch.cat([patches, size_embed, time_tokens], dim=1)
        # context slots stay in the token buffer between blocks (JiTBlock.forward: ctx_tail) when nothing upstream trains
        keep_slots = not cfg.do_context_fuse and not context_embed.requires_grad
        fresh = False
        for i, block in enumerate(self.blocks):
            with_ctx = i == cfg.context_start_block or (not cfg.do_context_fuse and i >= cfg.context_start_block)
            if with_ctx and tokens.shape[1] == pre_ctx:
                tokens = torch.cat([tokens, context_embed], dim=1)
            L = tokens.shape[1]
            has_ctx = L > pre_ctx
            refresh = keep_slots and has_ctx and i >= cfg.context_start_block
            tokens = self.forward_block(block, tokens, cos_sin[:L], seq_ctx if has_ctx else None,
                                        context_embed if refresh else None, pre_ctx, fresh)
            fresh = refresh
            if not cfg.do_context_fuse and i >= cfg.context_start_block and not keep_slots:
                tokens = tokens[:, :-ctx_len, :]
        ops.PREFETCH.reset()
        patches = self.final_layer(ops.token_prefix(tokens, n_patch) if tokens.is_cuda else tokens[:, :n_patch, :])
        return self.unpatchify(patches, height=height, width=width)


class Denoiser(JiT):
    pass
