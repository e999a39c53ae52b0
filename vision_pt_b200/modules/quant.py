"""NF4 quantised Linear behind the reference's quant-Linear registry.

Mirrors /root/reference/src/modules/quant/functional.py (QUANT_TYPE, validate_quant_type, _get_quant_linear,
replace_to_quant_linear, quantize_inplace, replace_by_prequantized_weights, quantize_state_dict) and
/root/reference/src/modules/quant/bnb.py:37-129 (BnbLinear4bit) for quant type ``bnb_nf4``.  The bitsandbytes calls
behind those (quantize_4bit, dequantize_4bit, MatMul4Bit) are replaced by libvptb200.so kernels; the state-dict key
set is the bitsandbytes one so checkpoints interchange.  Other quant types of the reference are out of scope and
raise NotImplementedError.
"""
from __future__ import annotations

import json
from typing import Literal

import torch
import torch.nn as nn

from .. import ops
from .state_dict import get_target_keys

QUANT_TYPE = Literal["fp8_e4m3fn", "bnb_int8", "bnb_fp4", "bnb_nf4", "quanto_int4", "quanto_int8", "ao_nf4", "ao_fp8"]
_ALL_TYPES = ["fp8_e4m3fn", "bnb_int8", "bnb_fp4", "bnb_nf4", "quanto_int4", "quanto_int8", "ao_nf4", "ao_fp8"]

BLOCKSIZE = 64
NESTED_BLOCKSIZE = 256

NF4_CODE = [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
            -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
            0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0]


def nested_code_table() -> torch.Tensor:
    """bitsandbytes create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): code of the nested absmax."""
    vals: list[float] = []
    for i in range(7):
        edges = torch.linspace(0.1, 1, 2 ** i + 1)
        mids = (edges[:-1] + edges[1:]) / 2.0
        vals += (10 ** (i - 6) * mids).tolist()
        vals += (-(10 ** (i - 6)) * mids).tolist()
    vals += [0, 1.0]
    vals.sort()
    return torch.tensor(vals, dtype=torch.float32)


def validate_quant_type(quant_type: str) -> None:
    if quant_type not in _ALL_TYPES:
        raise ValueError(f"Unknown quant_type: {quant_type}")


def _pack_meta(meta: dict) -> torch.Tensor:
    return torch.tensor(list(json.dumps(meta).encode("utf-8")), dtype=torch.uint8)


def _unpack_meta(t: torch.Tensor) -> dict:
    return json.loads(bytes(t.cpu().tolist()).decode("utf-8"))


class NF4Linear(nn.Linear):
    """Drop-in for ``BnbLinear4bit(quant_type="nf4")``: an nn.Linear whose weight is stored as bitsandbytes NF4 tensors.

    * constructible on ``meta``; ``weight`` is a frozen uint8 Parameter ``[(N*K+1)//2, 1]`` once quantised
    * ``_load_from_state_dict`` takes either a full-precision ``weight`` (quantised when the module reaches a CUDA
      device) or the prequantised key set ``weight``, ``weight.absmax``, ``weight.quant_map``, ``weight.nested_absmax``,
      ``weight.nested_quant_map``, ``weight.quant_state.bitsandbytes__nf4``
    * ``state_dict()`` writes the same keys back
    * ``forward`` = dequantise-in-the-GEMM-prologue on sm_100a; gradients flow to the input only
    """

    def __init__(self, input_features: int, output_features: int, bias: bool = True, compute_dtype=None,
                 compress_statistics: bool = True, quant_type: str = "nf4", quant_storage=torch.uint8, device=None):
        if quant_type != "nf4":
            raise NotImplementedError("only the bnb_nf4 path is built (fp4 is not on the hot path)")
        nn.Module.__init__(self)
        self.in_features = input_features
        self.out_features = output_features
        self.weight = nn.Parameter(torch.empty(output_features, input_features, dtype=compute_dtype, device="meta"),
                                   requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(output_features, dtype=compute_dtype, device="meta"), requires_grad=False)
        else:
            self.register_parameter("bias", None)
        self.compute_dtype = compute_dtype
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_storage = quant_storage
        self.quant_state: ops.Nf4Tensors | None = None

    # ------------------------------------------------------------------ quantisation / device moves
    @property
    def is_quantized(self) -> bool:
        return self.quant_state is not None

    def _quantize_now(self, device) -> None:
        w = self.weight.data.to(device)
        st = ops.nf4_quantize(w, nested_code_table(), torch.tensor(NF4_CODE, dtype=torch.float32))
        self.quant_state = st
        self.weight = nn.Parameter(st.packed, requires_grad=False)

    def _apply(self, fn, recurse=True):
        """`.to()` / `.cuda()`: a full-precision weight that reaches a CUDA device is quantised there
        (Params4bit behaviour); quantised tensors only change device, the bias follows dtype casts."""
        here = torch.device("cpu") if self.weight.is_meta else self.weight.device
        probe = fn(torch.empty(0, dtype=torch.float32, device=here))
        if self.quant_state is None:
            if probe.is_cuda and not self.weight.is_meta and self.weight.dtype.is_floating_point:
                if self.bias is not None and not self.bias.is_meta:
                    self.bias = nn.Parameter(fn(self.bias.data), requires_grad=False)
                self._quantize_now(probe.device)
                return self
            return super()._apply(fn, recurse)
        self.quant_state = self.quant_state.to(probe.device)
        self.weight = nn.Parameter(self.quant_state.packed, requires_grad=False)
        if self.bias is not None and not self.bias.is_meta:
            self.bias = nn.Parameter(fn(self.bias.data), requires_grad=False)
        return self

    # ------------------------------------------------------------------ (de)serialisation, bitsandbytes key set
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        st = self.quant_state
        if st is None:
            return
        meta = {"quant_type": "nf4", "blocksize": BLOCKSIZE, "dtype": str(st.dtype).replace("torch.", ""),
                "shape": list(st.shape), "nested_blocksize": NESTED_BLOCKSIZE, "nested_dtype": "float32",
                "nested_offset": st.offset}
        destination[prefix + "weight.absmax"] = st.absmax
        destination[prefix + "weight.quant_map"] = st.code
        destination[prefix + "weight.nested_absmax"] = st.nested_absmax
        destination[prefix + "weight.nested_quant_map"] = st.nested_code
        destination[prefix + "weight.quant_state.bitsandbytes__nf4"] = _pack_meta(meta)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        stats = {k[len(prefix) + len("weight."):]: v for k, v in state_dict.items() if k.startswith(prefix + "weight.")}
        if stats:
            tag = [k for k in stats if k.startswith("quant_state.bitsandbytes__")]
            if not tag:
                raise ValueError("quant_type not found")
            kind = tag[0][len("quant_state.bitsandbytes__"):]
            if kind != "nf4":
                raise NotImplementedError(f"prequantised {kind} weights are not on the hot path")
            meta = _unpack_meta(stats[tag[0]])
            # a module that already lives on a CUDA device keeps its tensors there (a CPU checkpoint loaded into a GPU
            # model must not leave host pointers behind for the kernels to dereference)
            here = None if self.weight.is_meta else self.weight.device
            if here is None and self.bias is not None and not self.bias.is_meta:
                here = self.bias.device
            mv = (lambda t: t.to(here)) if here is not None and here.type == "cuda" else (lambda t: t)
            packed = mv(state_dict[prefix + "weight"])
            self.quant_state = ops.Nf4Tensors(
                packed=packed, absmax=mv(stats["absmax"]), nested_absmax=mv(stats["nested_absmax"].float()),
                nested_code=mv(stats["nested_quant_map"].float()), code=mv(stats["quant_map"].float()),
                offset=float(meta["nested_offset"]), shape=(int(meta["shape"][0]), int(meta["shape"][1])),
                dtype=getattr(torch, meta["dtype"]))
            self.weight = nn.Parameter(packed, requires_grad=False)
            if self.bias is not None:
                self.bias = nn.Parameter(mv(state_dict[prefix + "bias"]), requires_grad=False)
            return
        # full-precision weights: plain nn.Linear loading; quantised on the move to CUDA (or right away if already there)
        w = state_dict.get(prefix + "weight")
        if w is None:
            missing_keys.append(prefix + "weight")
            return
        self.weight = nn.Parameter(w.detach(), requires_grad=False)
        self.quant_state = None
        if self.bias is not None:
            b = state_dict.get(prefix + "bias")
            if b is None:
                missing_keys.append(prefix + "bias")
            else:
                self.bias = nn.Parameter(b.detach(), requires_grad=False)
        if w.is_cuda:
            self._quantize_now(w.device)

    # ------------------------------------------------------------------ compute
    def dequantize(self, dtype: torch.dtype | None = None) -> torch.Tensor:
        if self.quant_state is None:
            raise RuntimeError("NF4Linear holds no quantised weight yet (move it to a CUDA device first)")
        return ops.nf4_dequantize(self.quant_state, dtype)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.quant_state is None:
            raise RuntimeError("NF4Linear.forward before quantisation: load weights and move the module to CUDA")
        return ops.nf4_lora_linear(x, self.quant_state, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, quant=nf4"


def _get_quant_linear(module: nn.Module, quant_type: QUANT_TYPE) -> nn.Module:
    if quant_type == "bnb_nf4":
        return NF4Linear(module.in_features, module.out_features, bias=module.bias is not None, quant_type="nf4")
    if quant_type in _ALL_TYPES:
        raise NotImplementedError(f"quant_type {quant_type} is outside the B200 hot path (only bnb_nf4 is built)")
    raise ValueError(f"Unknown quant_type: {quant_type}")


def _walk_linears(module: nn.Module, prefix: str = ""):
    for name, layer in module.named_children():
        full = f"{prefix}{name}"
        if isinstance(layer, nn.Linear):
            yield module, name, full, layer
        else:
            yield from _walk_linears(layer, f"{full}.")


def replace_to_quant_linear(model: nn.Module, quant_type: QUANT_TYPE, include_keys: list[str],
                            exclude_keys: list[str] = []) -> nn.Module:
    """Swap the selected nn.Linear modules for (still empty) quantised ones, before loading a checkpoint."""
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))
    for parent, name, full, layer in list(_walk_linears(model)):
        if full in targets:
            q = _get_quant_linear(layer, quant_type)
            q.requires_grad_(False)
            setattr(parent, name, q)
    return model


def quantize_inplace(model: nn.Module, quant_type: QUANT_TYPE, include_keys: list[str], exclude_keys: list[str] = []) -> None:
    """Swap + ``load_state_dict(assign=True)`` of the existing weights (quantised when they reach the GPU)."""
    validate_quant_type(quant_type)
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))
    for parent, name, full, layer in list(_walk_linears(model)):
        if full in targets and not isinstance(layer, NF4Linear):
            q = _get_quant_linear(layer, quant_type)
            q.load_state_dict(layer.state_dict(), assign=True)
            setattr(parent, name, q)


def get_quant_type_from_children_dict(children: dict[str, torch.Tensor]) -> QUANT_TYPE:
    for key in children:
        if "quant_state" in key:
            kind = key[len("quant_state.bitsandbytes__"):]
            if kind == "nf4":
                return "bnb_nf4"
            if kind == "fp4":
                return "bnb_fp4"
        elif "weight_format" in key:
            return "bnb_int8"
    raise ValueError("quant_type not found")


def replace_by_prequantized_weights(model: nn.Module, state_dict: dict[str, torch.Tensor]) -> None:
    """Sniff ``<name>.weight.<stat>`` keys and swap those Linears for quantised modules."""
    for parent, name, full, layer in list(_walk_linears(model)):
        children = {k[len(full) + len(".weight."):]: v for k, v in state_dict.items() if k.startswith(f"{full}.weight.")}
        if not children:
            continue
        q = _get_quant_linear(layer, get_quant_type_from_children_dict(children))
        q.requires_grad_(False)
        setattr(parent, name, q)


def quantize_state_dict(state_dict: dict[str, torch.Tensor], quant_type: QUANT_TYPE, include_keys: list[str],
                        exclude_keys: list[str] = []) -> dict[str, torch.Tensor]:
    """In-place NF4 quantisation of the selected tensors of a state dict (bitsandbytes packed key set, CPU tensors)."""
    if quant_type not in ("bnb_nf4",):
        raise NotImplementedError("Only bitsandbytes 4bit (nf4) quantization is supported")
    targets = set(get_target_keys(include_keys, exclude_keys, list(state_dict.keys())))
    ncode = nested_code_table()
    code = torch.tensor(NF4_CODE, dtype=torch.float32)
    for key in list(state_dict.keys()):
        if key not in targets:
            continue
        w = state_dict[key]
        st = ops.nf4_quantize(w.cuda(), ncode, code)
        state_dict[key] = st.packed.cpu()
        meta = {"quant_type": "nf4", "blocksize": BLOCKSIZE, "dtype": str(w.dtype).replace("torch.", ""),
                "shape": list(w.shape), "nested_blocksize": NESTED_BLOCKSIZE, "nested_dtype": "float32",
                "nested_offset": st.offset}
        state_dict[f"{key}.absmax"] = st.absmax.cpu()
        state_dict[f"{key}.quant_map"] = st.code.cpu()
        state_dict[f"{key}.nested_absmax"] = st.nested_absmax.cpu()
        state_dict[f"{key}.nested_quant_map"] = st.nested_code.cpu()
        state_dict[f"{key}.quant_state.bitsandbytes__nf4"] = _pack_meta(meta)
    return state_dict
