"""JiT extension blocks on the sm_100a kernels: PoPE, U-JiT skip-merge blocks, cross-attention JiT
blocks and TREAD token routing.  Mirrors /root/reference/src/models/jit/extension/{pope,uvit,cross}.py and
train/jit/class_to_image_tread.py:73-118."""
from .cross import CrossAttention, CrossJiTBlock  # noqa: F401
from .pope import PopeAttention, PopeEmbedder, apply_pope, pope_table  # noqa: F401
from .tread import keep_and_route_tokens, merge_routed_tokens  # noqa: F401
from .uvit import UJiTBlock  # noqa: F401
