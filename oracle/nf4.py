"""CPU oracle for the NF4 (QLoRA) weight format on the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this package; the product (``vision_pt_b200``) never does and fails loudly without its CUDA library.

PARITY UNPINNED for this file: the arithmetic lives in bitsandbytes 0.48.2 (``/root/reference/uv.lock:308-309``),
a third-party wheel that is neither vendored under /root/reference nor installable here or on the GPU box, and the
reference's own tests hold no golden vectors for it (tests/test_modules_quant.py:129-193 only check dtype, key names
and self-consistency).  What follows restates the published bitsandbytes algorithm
(``bitsandbytes/functional.py``: quantize_4bit / dequantize_4bit / quantize_blockwise / create_dynamic_map /
QuantState.as_dict and ``csrc/kernels.cu``: kQuantizeBlockwise, kDequantizeBlockwise, dQuantizeNF4, dQuantize)
as called from the reference at src/modules/quant/bnb.py:37-129 and src/modules/quant/functional.py:342-371.
The NF4 code table was re-derived with scipy (tests/test_oracle_golden.py: create_normal_map) and matches bit-for-bit.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np
import torch

BLOCK = 64          # quantize_4bit(blocksize=64)
NESTED_BLOCK = 256  # quantize_blockwise(absmax - offset, blocksize=256)

NF4_CODE = np.array(
    [-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
     -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
     0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0],
    dtype=np.float32)

# dQuantizeNF4 decision tree: strict ">" against these thresholds (midpoints of neighbouring codes), fp32 compare.
_NF4_THRESHOLDS = np.array(
    [-0.8480964004993439, -0.6106329262256622, -0.4599952697753906, -0.33967943489551544, -0.23460740596055984,
     -0.13791173323988914, -0.045525018125772476, 0.03979014977812767, 0.1202552504837513, 0.2035212516784668,
     0.2920137718319893, 0.3893125355243683, 0.5016634166240692, 0.6427869200706482, 0.8614784181118011],
    dtype=np.float32)


def dynamic_map_signed8() -> np.ndarray:
    """create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8): the 256-entry code for nested absmax."""
    data: list[float] = []
    for i in range(7):
        n = 2 ** i + 1
        bounds = torch.linspace(0.1, 1, n)
        means = (bounds[:-1] + bounds[1:]) / 2.0
        scale = 10 ** (-6 + i)
        data += (scale * means).tolist()
        data += (-scale * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 256
    data.sort()
    return torch.tensor(data, dtype=torch.float32).numpy()


def _nearest_code_8bit(code: np.ndarray, x: np.ndarray) -> np.ndarray:
    """dQuantize<0>: 7-step binary search from pivot 127 plus the midpoint rule (csrc/kernels.cu)."""
    x = x.astype(np.float32)
    n = x.shape[0]
    pivot = np.full(n, 127, np.int64)
    upper_p = np.full(n, 255, np.int64)
    lower_p = np.zeros(n, np.int64)
    lower = np.full(n, -1.0, np.float32)
    upper = np.full(n, 1.0, np.float32)
    val = code[pivot]
    i = 64
    while i > 0:
        gt = x > val
        lower_p = np.where(gt, pivot, lower_p)
        lower = np.where(gt, val, lower)
        upper_p = np.where(gt, upper_p, pivot)
        upper = np.where(gt, upper, val)
        pivot = np.where(gt, pivot + i, pivot - i)
        val = code[pivot]
        i >>= 1
    upper = np.where(upper_p == 255, code[255], upper)
    lower = np.where(lower_p == 0, code[0], lower)
    gt = x > val
    mid_up = ((upper + val) * np.float32(0.5)).astype(np.float32)
    mid_lo = ((lower + val) * np.float32(0.5)).astype(np.float32)
    res_gt = np.where(x > mid_up, upper_p, pivot)
    res_le = np.where(x < mid_lo, lower_p, pivot)
    return np.where(gt, res_gt, res_le).astype(np.uint8)


@dataclass
class Nf4State:
    """The tensors bitsandbytes keeps for one ``Params4bit`` with compress_statistics=True."""
    packed: torch.Tensor          # uint8 [(n+1)//2, 1]
    absmax: torch.Tensor          # uint8 [n/64]          codes of (absmax - offset)
    nested_absmax: torch.Tensor   # fp32  [ceil(n/64/256)]
    nested_code: torch.Tensor     # fp32  [256]
    code: torch.Tensor            # fp32  [16]
    offset: float
    shape: tuple[int, int]
    dtype: torch.dtype            # dtype of the tensor at quantisation time == dequantisation target

    def as_dict(self) -> dict[str, torch.Tensor]:
        """QuantState.as_dict(packed=True) key set, as written by quantize_state_dict (functional.py:361-368)."""
        meta = {
            "quant_type": "nf4", "blocksize": BLOCK, "dtype": str(self.dtype).replace("torch.", ""),
            "shape": list(self.shape), "nested_blocksize": NESTED_BLOCK, "nested_dtype": "float32",
            "nested_offset": self.offset,
        }
        blob = torch.tensor(list(json.dumps(meta).encode("utf-8")), dtype=torch.uint8)
        return {
            "absmax": self.absmax, "quant_map": self.code, "nested_absmax": self.nested_absmax,
            "nested_quant_map": self.nested_code, "quant_state.bitsandbytes__nf4": blob,
        }

    @staticmethod
    def from_dict(packed: torch.Tensor, stats: dict[str, torch.Tensor]) -> "Nf4State":
        meta = json.loads(bytes(stats["quant_state.bitsandbytes__nf4"].tolist()).decode("utf-8"))
        return Nf4State(packed=packed, absmax=stats["absmax"], nested_absmax=stats["nested_absmax"],
                        nested_code=stats["nested_quant_map"], code=stats["quant_map"],
                        offset=float(meta["nested_offset"]), shape=tuple(meta["shape"]),
                        dtype=getattr(torch, meta["dtype"]))


def quantize_nf4(w: torch.Tensor) -> Nf4State:
    """quantize_4bit(A, blocksize=64, compress_statistics=True, quant_type="nf4", quant_storage=uint8)."""
    assert w.dim() == 2
    flat = w.detach().reshape(-1).to(torch.float32).numpy()
    n = flat.shape[0]
    assert n % BLOCK == 0, "the reference shapes all give whole 64-element blocks over the flattened weight"
    blocks = flat.reshape(-1, BLOCK)
    absmax = np.abs(blocks).max(axis=1).astype(np.float32)
    inv = (np.float32(1.0) / absmax).astype(np.float32)
    x = (blocks * inv[:, None]).astype(np.float32)
    codes = (x[..., None] > _NF4_THRESHOLDS).sum(axis=-1).astype(np.uint8).reshape(-1)
    packed = (codes[0::2] << 4) | codes[1::2]
    # double quantisation of the statistics
    offset = np.float32(torch.from_numpy(absmax).mean().item())
    shifted = (absmax - offset).astype(np.float32)
    ncode = dynamic_map_signed8()
    nb = shifted.shape[0]
    pad = (-nb) % NESTED_BLOCK
    sp = np.concatenate([shifted, np.zeros(pad, np.float32)]).reshape(-1, NESTED_BLOCK)
    nested_absmax = np.abs(sp).max(axis=1).astype(np.float32)
    ninv = (np.float32(1.0) / nested_absmax).astype(np.float32)
    qa = _nearest_code_8bit(ncode, (sp * ninv[:, None]).astype(np.float32).reshape(-1))[:nb]
    return Nf4State(
        packed=torch.from_numpy(packed.copy()).reshape(-1, 1), absmax=torch.from_numpy(qa.copy()),
        nested_absmax=torch.from_numpy(nested_absmax.copy()), nested_code=torch.from_numpy(ncode.copy()),
        code=torch.from_numpy(NF4_CODE.copy()), offset=float(offset), shape=tuple(w.shape), dtype=w.dtype)


def dequantize_absmax(st: Nf4State) -> torch.Tensor:
    """dequantize_blockwise(absmax, state2) then ``absmax += offset``: two separately rounded fp32 operations."""
    qa = st.absmax.to(torch.int64)
    idx = torch.arange(qa.shape[0], device=qa.device) // NESTED_BLOCK        # runs wherever the tensors live
    am = st.nested_code.to(torch.float32)[qa] * st.nested_absmax.to(torch.float32)[idx]
    return am + torch.tensor(st.offset, dtype=torch.float32, device=qa.device)


def dequantize_nf4(st: Nf4State) -> torch.Tensor:
    """dequantize_4bit: w[i] = T_rn(fl32(code[nibble_i] * absmax[i // 64])), high nibble = even element."""
    am = dequantize_absmax(st)
    b = st.packed.reshape(-1).to(torch.int64)
    codes = torch.stack([b >> 4, b & 15], dim=1).reshape(-1)
    n = st.shape[0] * st.shape[1]
    vals = st.code.to(torch.float32)[codes[:n]] * am.repeat_interleave(BLOCK)[:n]
    return vals.to(st.dtype).reshape(st.shape)


def linear_nf4(x: torch.Tensor, st: Nf4State, bias: torch.Tensor | None) -> torch.Tensor:
    """MatMul4Bit.forward: F.linear(x, dequantize_4bit(W).to(x.dtype), bias); autograd gives grad_x = grad_y @ W."""
    w = dequantize_nf4(st).to(x.dtype)
    return torch.nn.functional.linear(x, w, None if bias is None else bias.to(x.dtype))
