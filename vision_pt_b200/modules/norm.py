"""Norm layers of the hot path -- mirror of /root/reference/src/modules/norm.py (FP32LayerNorm 9-17, FP32RMSNorm 20-27,
SingleAdaLayerNormZero 37-90, get_norm_layer 179-215) on fused sm_100a kernels.  DyT / Derf are out of scope."""
from __future__ import annotations

from typing import Literal, NamedTuple

import torch
import torch.nn as nn

from .. import ops

NormType = Literal["rms", "layer", "dyt", "derf"]


class FP32RMSNorm(nn.RMSNorm):
    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        eps = self.eps if self.eps is not None else torch.finfo(torch.float32).eps
        return ops.rms_norm(hidden_states, self.weight, eps)


class FP32LayerNorm(nn.LayerNorm):
    """LayerNorm with fp32 statistics over the last dimension, with or without affine parameters (reference norm.py:9-17:
    F.layer_norm on the fp32 copy, cast back): one fused pass, rows up to 4096 wide."""

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(hidden_states, self.weight, self.bias, self.eps)


class SingleAdaLayerNormZeroOutput(NamedTuple):
    hidden_states: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor
    gate: torch.Tensor


class SingleAdaLayerNormZero(nn.Module):
    def __init__(self, hidden_dim: int, gate_dim: int, embedding_dim: int) -> None:
        super().__init__()
        self.act = nn.SiLU()
        self.norm = FP32LayerNorm(hidden_dim, elementwise_affine=False, eps=1e-6)
        self.scale_shift = nn.Linear(embedding_dim, 2 * hidden_dim, bias=True)
        self.gate = nn.Linear(embedding_dim, gate_dim, bias=True)

    def init_weights(self) -> None:
        for lin in (self.scale_shift, self.gate):
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)

    def forward(self, hidden_states: torch.Tensor, time_embed: torch.Tensor) -> SingleAdaLayerNormZeroOutput:
        time_embed = self.act(time_embed)
        scale, shift = self.scale_shift(time_embed).chunk(2, dim=1)
        gate = self.gate(time_embed)
        out = ops.ln_modulate(hidden_states, scale, shift, self.norm.eps)   # LN(x) * (1 + scale) + shift, one pass
        return SingleAdaLayerNormZeroOutput(hidden_states=out, scale=scale, shift=shift, gate=gate)


def adaln_modulate(hidden_states: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """AdaLayerNormZero modulate of the CogView4 block (src/models/cogview4/denoiser.py:182-187)."""
    return ops.ln_modulate(hidden_states, scale, shift, eps)


def adaln_gate_residual(hidden_states: torch.Tensor, branch: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """hidden_states + branch * gate.unsqueeze(1) (src/models/cogview4/denoiser.py:401-420)."""
    return ops.gate_residual(hidden_states, branch, gate)


def get_norm_layer(norm_type: NormType, normalized_shape: int, elementwise_affine: bool = True, eps: float = 1e-6,
                   **kwargs) -> nn.Module:
    if norm_type == "rms":
        return FP32RMSNorm(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)
    if norm_type == "layer":
        return FP32LayerNorm(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)
    if norm_type in ("dyt", "derf"):
        raise NotImplementedError(f"norm_type {norm_type} is outside the B200 hot path")
    raise ValueError(f"Unknown norm_type: {norm_type}")
