from .denoiser import CrossAttention, FeedForward, GeGLU, SelfAttention, TransformerBlock  # noqa: F401
