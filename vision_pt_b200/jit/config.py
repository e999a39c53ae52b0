"""Mirror of /root/reference/src/models/jit/config.py:16-65 (field names and defaults kept verbatim: boundary)."""
from __future__ import annotations

from typing import Literal

from pydantic import BaseModel

NormType = Literal["rms", "layer", "dyt", "derf"]
PositionalEncoding = Literal["rope", "pope", "n-pope"]


class DenoiserConfig(BaseModel):
    patch_size: int = 16
    in_channels: int = 3
    out_channels: int = 3
    hidden_size: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: float = 4.0
    attn_dropout: float = 0.0
    proj_dropout: float = 0.0

    bottleneck_dim: int = 128
    use_output_bottleneck: bool = False
    use_pixel_shuffle: bool = False

    norm_type: NormType = "rms"

    num_time_tokens: int = 4
    timestep_scale: float = 1.0

    positional_encoding: PositionalEncoding = "rope"
    rope_theta: float = 256.0
    rope_axes_dims: list[int] = [16, 24, 24]
    rope_axes_lens: list[int] = [256, 128, 128]
    rope_zero_centered: list[bool] = [False, True, True]
    rope_do_normalize: list[bool] = [False, True, True]
    rope_normalize_by: float = 64.0

    context_dim: int = 768
    context_start_block: int = 0
    do_context_fuse: bool = False


class JiT_B_16_Config(DenoiserConfig):
    patch_size: int = 16
    depth: int = 12
    hidden_size: int = 768
    num_heads: int = 12
    bottleneck_dim: int = 128
    context_dim: int = 768
    context_start_block: int = 4
    rope_axes_dims: list[int] = [16, 24, 24]
    rope_axes_lens: list[int] = [256, 128, 128]


class JiT_L_16_Config(DenoiserConfig):
    """Upstream LTH14/JiT "L" sizes on the generic DenoiserConfig (BASELINE.json configs[2])."""
    depth: int = 24
    hidden_size: int = 1024
    num_heads: int = 16


class JiT_H_16_Config(DenoiserConfig):
    """Upstream "H" sizes (head_dim 80; BASELINE.json configs[3])."""
    depth: int = 32
    hidden_size: int = 1280
    num_heads: int = 16
    rope_axes_dims: list[int] = [16, 32, 32]
