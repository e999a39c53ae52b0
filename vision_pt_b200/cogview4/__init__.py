from .denoiser import (AdaLayerNormZero, AdaLayerNormZeroOutput, FeedForward, FinalAdaLayerNorm, SelfAttention,  # noqa: F401
                       TransformerBlock, apply_rotary_emb)
