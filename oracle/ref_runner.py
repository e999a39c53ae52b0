"""Runs the REFERENCE's own JiT modules (staged in oracle/_ref, see oracle/make_ref.py) through one NF4-QLoRA training
step: the CPU baseline (`bench.py --impl reference`, fp32 on the host cores) and the "reference-equivalent GPU" line
(`bench.py`'s `reference_gpu` leg, bf16 eager on the same B200).  TEST INFRASTRUCTURE ONLY (see
oracle/nf4.py for the import rule) -- it is a baseline that is MEASURED, never a path the product takes.

What runs is the reference's code: `Denoiser` (src/models/jit/denoiser.py: RoPE tables rebuilt per forward, expanded bool
mask, torch.cat per block, F.scaled_dot_product_attention), `LoRALinear` (src/modules/peft/lora.py:92-104) installed by
the reference's `PeftTargetConfig.replace_to_peft_layer`, `FP32RMSNorm`, and the step of train/jit/class_to_image.py:
166-243 (scale_shift_sigmoid timesteps, noising, `treat_loss` with model_pred = "image") followed by
clip_grad_norm_ + torch.optim.AdamW (src/trainer/common.py:376-388, src/models/for_training.py:98-109).
The one restated piece is the NF4 base linear: `BnbLinear4bit` needs the bitsandbytes wheel, which exists neither in this
image nor on the GPU box, so `RestatedLinear4bit` stands in for it with bitsandbytes' MatMul4Bit semantics (dequantise
in forward AND in backward, oracle/nf4.py, plain torch ops on whatever device the tensors live).
"""
from __future__ import annotations

import re
import time
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import nf4 as on
from . import refimport

BLOCK_LINEARS = ("to_q", "to_k", "to_v", "to_o", "w_1", "w_2", "w_3")
# JiT-B/16 is the reference's own JiT_B_16_Config; it ships no L / H classes, so those are the upstream "L" / "H" sizes on its
# generic DenoiserConfig (what BASELINE.json configs[2] / [3] name)
MODEL_OVERRIDES = {"JiT-L/16": dict(depth=24, hidden_size=1024, num_heads=16),
                   "JiT-H/16": dict(depth=32, hidden_size=1280, num_heads=16, rope_axes_dims=[16, 32, 32])}


def reference_modules() -> SimpleNamespace:
    refimport.load()
    from src.models.jit import config as jcfg
    from src.models.jit.denoiser import Denoiser
    from src.modules.peft import PeftTargetConfig
    from src.modules.peft.lora import LoRAConfig, LoRALinear
    from src.utils.state_dict import RegexMatch
    return SimpleNamespace(config=jcfg, Denoiser=Denoiser, PeftTargetConfig=PeftTargetConfig, LoRAConfig=LoRAConfig,
                           LoRALinear=LoRALinear, RegexMatch=RegexMatch)


class _MatMul4Bit(torch.autograd.Function):
    """bitsandbytes MatMul4Bit: y = x @ dequant(W)^T + b; backward dequantises again; no weight gradient."""

    @staticmethod
    def forward(ctx, x, state, bias):
        ctx.state = state
        w = on.dequantize_nf4(state).to(x.dtype)
        return torch.nn.functional.linear(x, w, bias)

    @staticmethod
    def backward(ctx, dy):
        w = on.dequantize_nf4(ctx.state).to(dy.dtype)
        return dy @ w, None, None


class RestatedLinear4bit(nn.Linear):
    """Stand-in for BnbLinear4bit(quant_type="nf4") (src/modules/quant/bnb.py:37-129): an nn.Linear whose weight is the
    packed uint8 tensor and whose forward is MatMul4Bit."""

    def __init__(self, state: on.Nf4State, bias: torch.Tensor | None):
        nn.Module.__init__(self)
        self.out_features, self.in_features = state.shape
        self.state = state
        self.weight = nn.Parameter(state.packed, requires_grad=False)
        self.bias = None if bias is None else nn.Parameter(bias, requires_grad=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _MatMul4Bit.apply(x, self.state, self.bias)


_WRAPPED = re.compile(r"^(blocks\.\d+\.(?:attn|mlp)\.\w+)\.linear\.(weight|bias)$")


def plain_state_dict(lora_wrapped_state: dict) -> dict:
    """State dict of a LoRA-wrapped, NF4-quantised model (ours or the reference's) -> the keys of the plain Denoiser:
    `blocks.N.attn.to_q.linear.bias` -> `blocks.N.attn.to_q.bias`; adapter tensors, packed NF4 weights and their statistics
    are dropped (the block weights are installed from Nf4State objects)."""
    out = {}
    for k, v in lora_wrapped_state.items():
        if ".lora_" in k or k.endswith(".alpha") or ".weight." in k or v.dtype == torch.uint8:
            continue
        m = _WRAPPED.match(k)
        out[f"{m.group(1)}.{m.group(2)}" if m else k] = v.detach().float().cpu()
    return out


def state_to(st: on.Nf4State, device) -> on.Nf4State:
    return on.Nf4State(st.packed.to(device), st.absmax.to(device), st.nested_absmax.to(device), st.nested_code.to(device),
                       st.code.to(device), st.offset, st.shape, st.dtype)


def build_reference_jit(model: str = "JiT-B/16", rank: int = 16, alpha: float = 16.0, device="cpu", dtype=torch.float32,
                        seed: int = 42, nf4_states: dict | None = None, state_dict: dict | None = None):
    """The reference's Denoiser, random-init (`init_weights`) or loaded from `state_dict`, block linears as NF4
    (quantised here from their bf16 values, or taken from `nf4_states`: {module path: Nf4State}), LoRA on the block
    linears through the reference's own PEFT entry point."""
    R = reference_modules()
    if isinstance(model, dict):
        fields = set(R.config.DenoiserConfig.model_fields)
        cfg = R.config.DenoiserConfig(**{k: v for k, v in model.items() if k in fields})
    elif model == "JiT-B/16":
        cfg = R.config.JiT_B_16_Config()
    else:
        cfg = R.config.DenoiserConfig(**MODEL_OVERRIDES[model])
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = R.Denoiser(cfg)
    if hasattr(net, "init_weights"):
        net.init_weights()
    elif hasattr(net, "initialize_weights"):
        net.initialize_weights()
    if state_dict is not None:
        missing, unexpected = net.load_state_dict(state_dict, strict=False)
        left = [k for k in missing if not any(f".{n}." in k for n in BLOCK_LINEARS)]
        assert not left, f"reference Denoiser: parameters missing from the given state dict: {left[:5]}"
    net.to(dtype)
    net.requires_grad_(False)
    for bi, blk in enumerate(net.blocks):
        for parent, names in ((blk.attn, BLOCK_LINEARS[:4]), (blk.mlp, BLOCK_LINEARS[4:])):
            for n in names:
                lin = getattr(parent, n)
                path = f"blocks.{bi}.{'attn' if parent is blk.attn else 'mlp'}.{n}"
                st = nf4_states[path] if nf4_states is not None else on.quantize_nf4(lin.weight.detach().to(torch.bfloat16).cpu())
                bias = lin.bias.detach().to(dtype) if lin.bias is not None else None
                setattr(parent, n, RestatedLinear4bit(state_to(st, "cpu"), bias))
    R.PeftTargetConfig(include_keys=[R.RegexMatch(regex=r"blocks\.\d+\.(attn|mlp)\.")],
                       config=R.LoRAConfig(rank=rank, alpha=alpha, dtype=str(dtype).replace("torch.", ""))
                       ).replace_to_peft_layer(net)
    net.to(device)
    for m in net.modules():
        if isinstance(m, RestatedLinear4bit):
            m.state = state_to(m.state, device)
    for name, p in net.named_parameters():
        p.requires_grad_(("lora_down" in name or "lora_up" in name) and "alpha" not in name)
    torch.random.set_rng_state(rng)
    return net, cfg


def train_steps(net, cfg, batch: int, height: int, width: int, steps: int, warmup: int, device, dtype,
                loss_target: str = "image", max_tokens: int = 64, seed: int = 0, on_step=None) -> dict:
    """`warmup` + `steps` optimisation steps of train/jit/class_to_image.py on one synthetic batch; wall-clock per step
    (device-synchronised when on a GPU)."""
    dev = torch.device(device)
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, height, width, generator=g).to(torch.float16)
    n_labels = torch.randint(8, 41, (batch,), generator=g)
    mask = (torch.arange(max_tokens).unsqueeze(0) < n_labels.unsqueeze(1)).to(torch.int64)
    context = (torch.randn(batch, max_tokens, cfg.context_dim, generator=g) * 0.02) * mask.unsqueeze(-1)
    image, mask, context = image.to(dev), mask.to(dev), context.to(dev)
    size = torch.tensor([[height, width]], device=dev).repeat(batch, 1)
    leaves = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(leaves, lr=1e-4, weight_decay=0.01)
    net.train()
    sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
    times, loss = [], None
    for it in range(warmup + steps):
        sync()
        t0 = time.perf_counter()
        t = (torch.randn(batch, device=dev) * 0.8 - 0.8).sigmoid()
        noise = torch.randn_like(image)
        tv = t.view(-1, 1, 1, 1).to(image.dtype)
        noisy = tv * image + (1 - tv) * noise
        pred = net(image=noisy.to(dtype), timestep=t.to(dtype), context=context.to(dtype), context_mask=mask,
                   original_size=size, target_size=size, crop_coords=torch.zeros_like(size))
        if loss_target == "velocity":
            den = (1 - t.view(-1, 1, 1, 1)).clamp_min(0.05)
            loss = torch.nn.functional.mse_loss((pred - noisy) / den, (image - noisy) / den)
        else:
            loss = torch.nn.functional.mse_loss(pred, image.to(pred.dtype))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(leaves, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        sync()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if on_step is not None and on_step(it, dt) is False:
            break
    sec = sum(times) / max(1, len(times))
    return {"images_per_s": batch / sec if times else 0.0, "s_per_step": sec, "batch": batch, "steps": len(times),
            "warmup": warmup, "threads": torch.get_num_threads(), "loss": float(loss.detach()) if loss is not None else None}
