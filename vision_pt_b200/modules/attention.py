"""``scaled_dot_product_attention`` with a B200 backend.

Mirrors /root/reference/src/modules/attention.py:98-159 (argument names, dtype cast of fp32 inputs, error behaviour).
The ``"eager"`` / ``"sdpa"`` / ``"b200"`` backends run the sm_100a tcgen05 kernel; flash-attn / xformers are never
called.  Masks must be key-padding masks (True = attend, broadcast over heads and queries, valid keys first), which is
what JiT builds (src/models/jit/denoiser.py:375-381); they are passed to the kernel as per-sample key lengths.
"""
from __future__ import annotations

from typing import Literal

import torch

from .. import ops

AttentionImplementation = Literal["eager", "flash_attention_2", "xformers", "sdpa", "b200"]


def prefix_key_lengths(mask2d: torch.Tensor) -> torch.Tensor:
    """[B, Lk] key-padding mask (bool / integer, True = attend) -> int32 [B] number of leading valid keys.

    The kernels take key LENGTHS, so the valid keys of every sample must come first (what JiT's class / text encoders
    emit, reference class_encoder.py:72-81).  Anything else -- left padding, holes, a float additive mask -- is refused
    instead of being silently computed as a prefix: floating-point masks raise TypeError; the prefix property is checked
    on the host for CPU masks and by a device-side assert for CUDA masks (skipped only while a CUDA graph is being
    captured: the eager warm-up that precedes every capture has run the check on the same buffers)."""
    if mask2d.dtype.is_floating_point or mask2d.dtype.is_complex:
        raise TypeError("attention masks must be bool / integer key-padding masks (True = attend); additive float masks "
                        "are not supported by the B200 kernels")
    m = mask2d.to(torch.bool)
    lens = m.sum(dim=-1, dtype=torch.int32)
    is_prefix = (m == (torch.arange(m.shape[-1], device=m.device).unsqueeze(0) < lens.unsqueeze(1))).all()
    if not m.is_cuda:
        if not bool(is_prefix):
            raise NotImplementedError("only prefix key-padding masks (valid keys first) run on the B200 attention kernel")
    elif not torch.cuda.is_current_stream_capturing():
        torch._assert_async(is_prefix, "attention mask is not a prefix key-padding mask (valid keys must come first)")
    return lens.contiguous()


def key_lengths_from_mask(mask: torch.Tensor | None, batch: int, lk: int) -> torch.Tensor | None:
    """mask broadcastable to [B,H,Lq,Lk] that only depends on (b, key) -> int32 [B] valid-key counts."""
    if mask is None:
        return None
    m = mask
    if m.dim() == 2:
        m = m.view(batch, 1, 1, lk)
    if m.dim() != 4:
        raise ValueError("attention mask must be [B, Lk] or broadcastable to [B, H, Lq, Lk]")
    for d in (1, 2):
        if m.shape[d] != 1 and m.stride(d) != 0:
            raise NotImplementedError("only key-padding masks (constant over heads and queries) run on the B200 kernel")
    return prefix_key_lengths(m[:, 0, 0, :])


def scaled_dot_product_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor | None = None,
                                 scale: float | None = None, dropout: float = 0.0,
                                 backend: AttentionImplementation = "eager", attention_dtype: torch.dtype = torch.bfloat16,
                                 is_causal: bool = False) -> torch.Tensor:
    assert q.dim() == k.dim() == v.dim() == 4  # (batch_size, num_heads, seq_len, head_dim)
    if q.dtype == torch.float32:
        q, k, v = q.to(attention_dtype), k.to(attention_dtype), v.to(attention_dtype)
    if backend in ("flash_attention_2", "xformers"):
        if backend == "flash_attention_2" and mask is not None:
            raise ValueError("Flash Attention does not support attention masks")
        backend = "b200"  # same contract, served by the sm_100a kernel
    if backend not in ("eager", "sdpa", "b200"):
        raise ValueError(f"Unknown backend: {backend}")
    if is_causal or dropout != 0.0:
        raise NotImplementedError("the B200 attention kernel is non-causal and dropout-free (the JiT/DiT/SDXL setting)")
    if q.dtype != torch.bfloat16:
        raise TypeError("the B200 attention kernel runs in bfloat16")
    seqlens = key_lengths_from_mask(mask, q.shape[0], k.shape[2])
    return ops.attention(q, k, v, seqlens, scale)


def get_attn_implementation_label(use_flash_attention: bool) -> AttentionImplementation:
    return "b200"
