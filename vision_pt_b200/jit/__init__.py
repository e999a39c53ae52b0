"""JiT denoiser on the B200 kernels -- mirror of /root/reference/src/models/jit (config.py, denoiser.py)."""
from .config import DenoiserConfig, JiT_B_16_Config, JiT_H_16_Config, JiT_L_16_Config
from .denoiser import Denoiser, JiT, JiTBlock

__all__ = ["DenoiserConfig", "JiT_B_16_Config", "JiT_L_16_Config", "JiT_H_16_Config", "Denoiser", "JiT", "JiTBlock"]
