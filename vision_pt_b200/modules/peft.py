"""LoRA wrapper behind the reference's PEFT API.

Mirrors /root/reference/src/modules/peft: ``PeftLayer`` (util.py:10-49), ``LoRAConfig`` / ``LoRALinear``
(lora.py:11-163), ``replace_to_peft_layer`` / ``get_adapter_parameters`` / ``load_peft_weight`` /
``while_peft_disabled`` (functional.py:59-360) and ``PeftTargetConfig`` (__init__.py:24-48).  Adapter parameter names
(``lora_down.weight``, ``lora_up.weight``, ``alpha``) are the reference's, so saved adapters interchange.
When the wrapped layer is an ``NF4Linear`` (or a bf16 ``nn.Linear``) on CUDA, base GEMM, LoRA branch, scale and add run
as ONE fused sm_100a kernel; LoRAConv2d / LoHa are out of scope.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Callable, Literal, NamedTuple

import torch
import torch.nn as nn
from pydantic import BaseModel, field_validator

from .. import ops
from .quant import NF4Linear
from .state_dict import RegexMatch, get_target_keys

PEFT_TYPE = Literal["lora", "loha", "none"]
_DTYPES = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}


class PeftConfigMixin(BaseModel):
    type: PEFT_TYPE
    dtype: str = "bfloat16"


class LoRAConfig(PeftConfigMixin):
    type: Literal["lora"] = "lora"
    rank: int
    alpha: float = 1.0
    dropout: float = 0.0
    use_bias: bool = False


class PeftLayer(ABC, nn.Module):
    adapter_param_names: list[str]
    adapter_weight_names: list[str]
    enabled: bool

    @abstractmethod
    def init_weights(self) -> None: ...

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = enabled

    @abstractmethod
    def forward(self, x: torch.Tensor) -> torch.Tensor: ...

    @classmethod
    @abstractmethod
    def from_weights(cls, adapter_weights: dict[str, torch.Tensor], original_layer: nn.Module) -> "PeftLayer": ...

    @abstractmethod
    def load_weights(self, adapter_weights: dict[str, torch.Tensor | None]) -> None: ...


class LoRALinear(PeftLayer):
    adapter_param_names = ["lora_up", "lora_down", "alpha"]
    adapter_weight_names = ["lora_up.weight", "lora_up.bias", "lora_down.weight", "alpha"]

    def __init__(self, config: LoRAConfig, original_linear: nn.Linear) -> None:
        super().__init__()
        self.config = config
        dtype = _DTYPES[config.dtype]
        k, n = original_linear.in_features, original_linear.out_features
        self.lora_down = nn.Linear(k, config.rank, bias=False, dtype=dtype)
        self.lora_up = nn.Linear(config.rank, n, bias=False, dtype=dtype)
        self.dropout = nn.Dropout(config.dropout) if config.dropout > 0 else nn.Identity()
        self.alpha = nn.Parameter(torch.tensor(config.alpha, dtype=dtype), requires_grad=False)
        self._alpha_value = float(config.alpha)   # host copy: the kernel scale never costs a device sync
        self.rank = config.rank
        if config.use_bias:
            self.lora_up.bias = nn.Parameter(torch.zeros(n, dtype=dtype))
        self.enabled = True
        self.linear = original_linear
        self.linear.weight.requires_grad_(False)
        if self.linear.bias is not None:
            self.linear.bias.requires_grad_(False)
        self.init_weights()

    def init_weights(self) -> None:
        dev = self.linear.weight.device
        if dev.type == "meta":
            dev = torch.device("cpu")
        self.lora_down.to_empty(device=dev)
        self.lora_up.to_empty(device=dev)
        nn.init.kaiming_uniform_(self.lora_down.weight)
        nn.init.zeros_(self.lora_up.weight)
        if self.lora_up.bias is not None:
            nn.init.zeros_(self.lora_up.bias)
        self.alpha = nn.Parameter(torch.tensor(self.config.alpha, dtype=self.lora_down.weight.dtype, device=dev),
                                  requires_grad=False)
        self._alpha_value = float(self.config.alpha)
        self.dropout = nn.Dropout(self.config.dropout) if self.config.dropout > 0 else nn.Identity()

    @property
    def scale(self) -> float:
        return self._alpha_value / self.rank

    def __setattr__(self, name, value):
        # `module.alpha = ...` keeps the host copy the fused kernels scale by in step with the stored parameter
        super().__setattr__(name, value)
        if name == "alpha" and isinstance(value, torch.Tensor) and not value.is_meta:
            object.__setattr__(self, "_alpha_value", float(value.detach().float().item()))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # load_state_dict() copies into `alpha` in place: refresh the host copy (reference lora.py:92-104 reads self.alpha)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        a = state_dict.get(prefix + "alpha")
        if a is not None and not a.is_meta:
            self._alpha_value = float(a.detach().float().item())

    @property
    def fusable(self) -> bool:
        """True when base GEMM + LoRA branch can run as the single fused kernel (CUDA bf16, rank <= 16, no dropout)."""
        base = self.linear
        plain = (type(base) is nn.Linear and base.weight.dtype == torch.bfloat16 and base.weight.is_cuda
                 and base.in_features % 8 == 0)
        quant = isinstance(base, NF4Linear) and base.is_quantized and base.weight.is_cuda
        return ((plain or quant) and isinstance(self.dropout, nn.Identity)
                and self.lora_up.bias is None and self.rank <= ops.RANK
                and self.lora_down.weight.dtype == torch.bfloat16 and self.lora_down.weight.is_cuda)

    def forward(self, x: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:
        if not self.enabled:
            out = self.linear(x)
            return out if residual is None else out + residual
        if x.is_cuda and self.fusable:
            base = self.linear
            w = base.quant_state if isinstance(base, NF4Linear) else base.weight
            return ops.nf4_lora_linear(x, w, base.bias, self.lora_down.weight, self.lora_up.weight, self.scale, residual)
        # composition of the reference (dropout, LoRA bias, rank > 16, fp32 adapters): same math, separate launches
        out = self.linear(x)
        lora = self.lora_up(self.lora_down(self.dropout(x)))
        out = out + lora * (self.alpha / self.rank)
        return out if residual is None else out + residual

    def train(self, mode: bool = True) -> "LoRALinear":
        self.lora_down.train(mode)
        self.lora_up.train(mode)
        self.linear.train(False)
        return self

    def requires_grad_(self, requires_grad: bool = True) -> "LoRALinear":
        self.lora_down.requires_grad_(requires_grad)
        self.lora_up.requires_grad_(requires_grad)
        self.linear.weight.requires_grad_(False)
        return self

    @classmethod
    def from_weights(cls, adapter_weights: dict[str, torch.Tensor], original_layer: nn.Linear) -> "LoRALinear":
        rank = adapter_weights["lora_down.weight"].shape[0]
        alpha = float(adapter_weights["alpha"].item())
        module = cls(LoRAConfig(rank=rank, alpha=alpha), original_layer)
        module.load_weights(adapter_weights)
        return module

    def load_weights(self, adapter_weights: dict[str, torch.Tensor | None]) -> None:
        dev = self.lora_down.weight.device
        if (w := adapter_weights.get("lora_down.weight")) is not None:
            self.lora_down.weight = nn.Parameter(w.to(dev))
        if (w := adapter_weights.get("lora_up.weight")) is not None:
            self.lora_up.weight = nn.Parameter(w.to(dev))
        if (w := adapter_weights.get("lora_up.bias")) is not None:
            self.lora_up.bias = nn.Parameter(w.to(dev))
        if (w := adapter_weights.get("alpha")) is not None:
            self.alpha = nn.Parameter(w.to(dev), requires_grad=False)
            self._alpha_value = float(w.item())


def _get_peft_linear(module: nn.Linear, config: PeftConfigMixin) -> PeftLayer:
    if config.type == "none":
        raise ValueError("peft type 'none' is not parameter efficient training")
    if config.type == "lora":
        return LoRALinear(config=LoRAConfig.model_validate(config.model_dump()), original_linear=module)
    if config.type == "loha":
        raise NotImplementedError("LoHa is outside the B200 hot path")
    raise ValueError(f"Unknown peft type: {config.type}")


def _replace_to_peft_layer(model: nn.Module, config: PeftConfigMixin, target_keys: set[str], prefix: str = "") -> None:
    for name, layer in model.named_children():
        full = f"{prefix}{name}"
        if isinstance(layer, PeftLayer):
            continue
        if isinstance(layer, nn.Linear):
            if full in target_keys:
                setattr(model, name, _get_peft_linear(layer, config))
        elif isinstance(layer, nn.Conv2d):
            if full in target_keys:
                raise NotImplementedError("LoRAConv2d is outside the B200 hot path")
        else:
            _replace_to_peft_layer(layer, config, target_keys, f"{full}.")


def replace_to_peft_layer(model: nn.Module, include_keys: list[str | RegexMatch], exclude_keys: list[str | RegexMatch],
                          config: PeftConfigMixin, freeze_base: bool = False) -> None:
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))
    if freeze_base:
        for _, module in model.named_modules():
            module.requires_grad_(False)
    _replace_to_peft_layer(model, config, targets)


def get_adapter_parameters(model: nn.Module) -> dict[str, torch.Tensor]:
    out: dict[str, torch.Tensor] = {}
    for name, module in model.named_modules():
        names = getattr(module, "adapter_param_names", None)
        if names is None:
            continue
        for key, value in module.state_dict().items():
            if any(key.startswith(n) for n in names):
                out[f"{name}.{key}".replace("_orig_mod.", "")] = value
    return out


def detect_peft_method(state_dict: dict[str, torch.Tensor]) -> PEFT_TYPE:
    return "lora" if any(k.endswith(".lora_up.weight") for k in state_dict) else "none"


def _load_peft_weight(model: nn.Module, state_dict: dict[str, torch.Tensor], prefix: str = "") -> None:
    for name, layer in model.named_children():
        full = f"{prefix}{name}"
        weights = {n: state_dict.get(f"{full}.{n}") for n in LoRALinear.adapter_weight_names}
        complete = all(v is not None for k, v in weights.items() if "bias" not in k)
        if isinstance(layer, PeftLayer):
            if complete:
                layer.load_weights(weights)
        elif isinstance(layer, nn.Linear):
            if complete:
                setattr(model, name, LoRALinear.from_weights(weights, layer))
        else:
            _load_peft_weight(layer, state_dict, f"{full}.")


def load_peft_weight(model: nn.Module, state_dict: dict[str, torch.Tensor]) -> None:
    if detect_peft_method(state_dict) == "none":
        raise ValueError("Failed to detect peft method from state_dict")
    _load_peft_weight(model, state_dict)


class TrainableParameters(NamedTuple):
    trainable_params: int
    all_param: int
    trainable_percent: float


def calculate_trainable_parameters(model: nn.Module) -> TrainableParameters:
    total = sum(p.numel() for p in model.parameters())
    train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    return TrainableParameters(train, total, 100 * train / max(total, 1))


def print_trainable_parameters(model: nn.Module, print_fn: Callable = print) -> None:
    t = calculate_trainable_parameters(model)
    print_fn(f"Trainable params: {t.trainable_params}, All params: {t.all_param}, Trainable%: {t.trainable_percent:.4f}%")
    if t.trainable_params == 0:
        warnings.warn("No trainable parameters found; check the peft config")


def set_peft_layer_enabled(model: nn.Module, enabled: bool) -> None:
    for _, module in model.named_modules():
        if hasattr(module, "set_enabled"):
            module.set_enabled(enabled)


@contextmanager
def while_peft_disabled(model: nn.Module):
    try:
        set_peft_layer_enabled(model, False)
        yield
    finally:
        set_peft_layer_enabled(model, True)


@contextmanager
def while_peft_enabled(model: nn.Module):
    try:
        set_peft_layer_enabled(model, True)
        yield
    finally:
        set_peft_layer_enabled(model, False)


class PeftTargetConfig(BaseModel):
    include_keys: list[str | RegexMatch] = []
    exclude_keys: list[str | RegexMatch] = []
    config: LoRAConfig
    resume_weight_path: str | None = None
    resume_rename_key_map: dict[str, str] = {}

    @field_validator("include_keys")
    def check_include_keys(cls, v):
        if len(v) == 0:
            raise ValueError("include_keys must not be empty")
        return v

    def replace_to_peft_layer(self, model: nn.Module, freeze_base: bool = False) -> None:
        replace_to_peft_layer(model, self.include_keys, self.exclude_keys, self.config, freeze_base=freeze_base)
