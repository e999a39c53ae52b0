"""CPU restatement of one JiT NF4-QLoRA training step, for timing the reference's CPU path.  TEST INFRASTRUCTURE ONLY
(see oracle/nf4.py for the import rule: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference).

What it follows (paths under /root/reference):
  train/jit/class_to_image.py:166-243   train_step: noising, denoiser call, treat_loss
  src/modules/quant/bnb.py:37-129       NF4 base linears (bitsandbytes MatMul4Bit: dequantise in forward AND in backward)
  src/modules/peft/lora.py:92-104       LoRA branch
  src/trainer/common.py:376-388         backward + optimizer step (torch.optim.AdamW over the LoRA matrices)
The forward is oracle.jit.jit_forward (pinned to the reference's own modules by tests/golden); the arithmetic type is
fp32 because that is what the reference runs on a CPU (BASELINE.json configs[0]).
"""
from __future__ import annotations

import math
import time

import torch

from . import jit as oj
from . import nf4 as on

BLOCK_LINEARS = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_o", "mlp.w_1", "mlp.w_2", "mlp.w_3")

JIT_CONFIGS = {
    "JiT-B/16": dict(patch_size=16, in_channels=3, out_channels=3, hidden_size=768, depth=12, num_heads=12, mlp_ratio=4.0,
                     bottleneck_dim=128, num_time_tokens=4, timestep_scale=1.0, rope_theta=256.0, rope_axes_dims=[16, 24, 24],
                     context_dim=768, context_start_block=4, do_context_fuse=False),
    "JiT-L/16": dict(patch_size=16, in_channels=3, out_channels=3, hidden_size=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                     bottleneck_dim=128, num_time_tokens=4, timestep_scale=1.0, rope_theta=256.0, rope_axes_dims=[16, 24, 24],
                     context_dim=768, context_start_block=0, do_context_fuse=False),
}


class _MatMul4Bit(torch.autograd.Function):
    """bitsandbytes MatMul4Bit: y = x @ dequant(W)^T + b; backward dequantises again, no weight gradient."""

    @staticmethod
    def forward(ctx, x, state, bias):
        ctx.state = state
        w = on.dequantize_nf4(state).to(x.dtype)
        return torch.nn.functional.linear(x, w, bias)

    @staticmethod
    def backward(ctx, dy):
        w = on.dequantize_nf4(ctx.state).to(dy.dtype)
        return dy @ w, None, None


def random_params(cfg: dict, rank: int = 16, nf4: bool = True, seed: int = 42, dtype=torch.float32) -> dict:
    """Random-init JiT parameter dict in the reference's LoRA-wrapped key layout (JiT.initialize_weights:
    Linear ~ N(0, 0.02^2), bias 0, norm weight 1); block linears quantised to NF4 from their bf16 values."""
    g = torch.Generator().manual_seed(seed)
    D, depth = cfg["hidden_size"], cfg["depth"]
    F_ = int(int(D * cfg["mlp_ratio"]) * 2 / 3)
    p, C, bd = cfg["patch_size"], cfg["in_channels"], cfg["bottleneck_dim"]
    rn = lambda *s: (torch.randn(*s, generator=g) * 0.02).to(dtype)
    P = {
        "patch_embedder.proj_1.weight": rn(bd, C, p, p), "patch_embedder.proj_2.weight": rn(D, bd, 1, 1),
        "patch_embedder.proj_2.bias": torch.zeros(D, dtype=dtype),
        "time_position_embeds": rn(cfg["num_time_tokens"], D),
        "context_embedder.weight": rn(D, cfg["context_dim"]), "context_embedder.bias": torch.zeros(D, dtype=dtype),
        "final_layer.norm_final.weight": torch.ones(D, dtype=dtype),
        "final_layer.linear.weight": rn(p * p * cfg["out_channels"], D),
        "final_layer.linear.bias": torch.zeros(p * p * cfg["out_channels"], dtype=dtype),
    }
    for emb in ("time_embedder", "image_size_embedder"):
        P[f"{emb}.mlp.0.weight"], P[f"{emb}.mlp.0.bias"] = rn(D, 256), torch.zeros(D, dtype=dtype)
        P[f"{emb}.mlp.2.weight"], P[f"{emb}.mlp.2.bias"] = rn(D, D), torch.zeros(D, dtype=dtype)
    for n, (k_, n_) in (("w_1", (D, F_)), ("w_2", (D, F_)), ("w_3", (F_, D))):
        P[f"final_layer.mlp.{n}.weight"], P[f"final_layer.mlp.{n}.bias"] = rn(n_, k_), torch.zeros(n_, dtype=dtype)
    shapes = {"attn.to_q": (D, D), "attn.to_k": (D, D), "attn.to_v": (D, D), "attn.to_o": (D, D), "mlp.w_1": (D, F_),
              "mlp.w_2": (D, F_), "mlp.w_3": (F_, D)}
    for i in range(depth):
        pre = f"blocks.{i}."
        for nm in ("norm1", "norm2"):
            P[f"{pre}{nm}.weight"] = torch.ones(D, dtype=dtype)
        for nm in ("attn.q_norm", "attn.k_norm"):
            P[f"{pre}{nm}.weight"] = torch.ones(D // cfg["num_heads"], dtype=dtype)
        for nm, (K, N) in shapes.items():
            w = rn(N, K)
            P[f"{pre}{nm}.linear.weight"] = on.quantize_nf4(w.to(torch.bfloat16)) if nf4 else w
            P[f"{pre}{nm}.linear.bias"] = torch.zeros(N, dtype=dtype)
            bound = 1.0 / math.sqrt(K)   # kaiming_uniform_(a=sqrt(5)) of nn.Linear / LoRALinear.init_weights
            P[f"{pre}{nm}.lora_down.weight"] = ((torch.rand(rank, K, generator=g) * 2 - 1) * bound).to(dtype).requires_grad_(True)
            P[f"{pre}{nm}.lora_up.weight"] = torch.zeros(N, rank, dtype=dtype).requires_grad_(True)
    return P


def make_batch(cfg: dict, batch: int, height: int, width: int, max_tokens: int = 64, seed: int = 0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, height, width, generator=g).to(torch.float16).to(dtype)
    n_labels = torch.randint(8, 41, (batch,), generator=g)
    mask = (torch.arange(max_tokens).unsqueeze(0) < n_labels.unsqueeze(1)).to(torch.int64)
    context = (torch.randn(batch, max_tokens, cfg["context_dim"], generator=g) * 0.02).to(dtype) * mask.unsqueeze(-1)
    return image, context, mask


def _patch_nf4(enable: bool):
    """Route NF4 weights through _MatMul4Bit (dequantise in forward and backward, like bitsandbytes)."""
    if not enable:
        return lambda: None
    orig = oj.lora_linear

    def lora_linear(x, w, bias, down=None, up=None, alpha: float = 1.0):
        if isinstance(w, on.Nf4State):
            out = _MatMul4Bit.apply(x, w, None if bias is None else bias.to(x.dtype))
            if down is None:
                return out
            t = torch.nn.functional.linear(x, down.to(x.dtype))
            u = torch.nn.functional.linear(t, up.to(x.dtype))
            return out + u * (alpha / down.shape[0])
        return orig(x, w, bias, down, up, alpha)

    oj.lora_linear = lora_linear
    return lambda: setattr(oj, "lora_linear", orig)


def time_train_steps(model: str = "JiT-B/16", batch: int = 4, height: int = 256, width: int = 256, steps: int = 2,
                     warmup: int = 1, rank: int = 16, alpha: float = 16.0, nf4: bool = True, threads: int | None = None,
                     loss_target: str = "image") -> dict:
    """Times `steps` CPU training steps (fwd + bwd + AdamW on the LoRA matrices).  Returns images/s and the setup."""
    if threads:
        torch.set_num_threads(threads)
    cfg = JIT_CONFIGS[model]
    P = random_params(cfg, rank=rank, nf4=nf4)
    leaves = [v for k, v in P.items() if isinstance(v, torch.Tensor) and v.requires_grad]
    opt = torch.optim.AdamW(leaves, lr=1e-4, weight_decay=0.01)
    image, context, mask = make_batch(cfg, batch, height, width)
    size = torch.tensor([[height, width]]).repeat(batch, 1)
    restore = _patch_nf4(nf4)
    times = []
    try:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            t = (torch.randn(batch) * 0.8 - 0.8).sigmoid()
            noise = torch.randn_like(image)
            tv = t.view(-1, 1, 1, 1)
            noisy = tv * image + (1 - tv) * noise
            pred = oj.jit_forward(P, cfg, noisy, t, context, size, size, torch.zeros_like(size), context_mask=mask, alpha=alpha)
            if loss_target == "velocity":
                loss = oj.velocity_loss(pred, image, noisy, t)
            else:
                loss = torch.nn.functional.mse_loss(pred, image)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(leaves, 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        restore()
    sec = sum(times) / len(times)
    return {"images_per_s": batch / sec, "s_per_step": sec, "batch": batch, "steps": steps, "threads": torch.get_num_threads(),
            "loss": float(loss.detach())}
