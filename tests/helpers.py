"""Shared test helpers: reference-name parameter dicts for the oracle, error metrics."""
import torch


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max |got - ref| / max |ref|  -- the block-output / gradient metric of BASELINE.json (<= 2e-2 in bf16)."""
    g, r = got.detach().float().cpu(), ref.detach().float().cpu()
    return float((g - r).abs().max() / r.abs().max().clamp_min(1e-12))


def lora_param_dict(base_state: dict, lora_state: dict) -> dict:
    """Oracle parameter dict for a LoRA-wrapped model: wrapped layers appear as `X.linear.weight` + `X.lora_*`."""
    wrapped = {k[: -len(".lora_down.weight")] for k in lora_state if k.endswith(".lora_down.weight")}
    out = {}
    for k, v in base_state.items():
        stem = k.rsplit(".", 1)[0]
        out[f"{stem}.linear.{k.rsplit('.', 1)[1]}" if stem in wrapped else k] = v
    out.update(lora_state)
    return out
