"""vision_pt_b200 -- B200-native (sm_100a) kernels for the JiT/DiT NF4-QLoRA block training hot path of
p1atdev/vision-pt, behind the reference's own module API (quant-Linear factory, PeftLayer/LoRALinear,
scaled_dot_product_attention, get_norm_layer, patchify).  Importing the package loads libvptb200.so; there is no
CPU or eager fallback."""
from . import _lib

_lib.load()

from . import ops  # noqa: E402
from .modules import attention, norm, patch, peft, quant  # noqa: E402

__all__ = ["ops", "attention", "norm", "patch", "peft", "quant"]
