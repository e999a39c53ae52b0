"""Torch-facing wrappers of the C ABI: raw launches (no autograd) and the autograd Functions built from them.

PyTorch only owns memory and streams here; every computation is a kernel of libvptb200.so.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import AttnTensorC, DequantItemC, LinearArgsC, LoraGradItemC, Nf4WeightC

RANK = 16  # LoRA rank of the fused kernels; smaller ranks are zero-padded

# How an NF4 weight reaches the tensor cores (include/vptb200.h: vpt_linear_args.w_scratch):
#   "scratch"  : dequantised once per call into an L2-resident bf16 workspace, main loop fed by TMA  (large M)
#   "prologue" : dequantised per pipeline stage by the producer warps of every CTA                    (small M)
#   "auto"     : scratch when M >= NF4_SCRATCH_MIN_M
NF4_GEMM_MODE = "auto"
NF4_SCRATCH_MIN_M = 1024
# SwiGLU forward / backward in the epilogues of the w_2 / w_3-backward GEMMs of the fused block (large-M route); off =
# separate swiglu_fwd / swiglu_bwd kernels (A/B measurements, VPT_FUSE_SWIGLU=0)
FUSE_SWIGLU = os.environ.get("VPT_FUSE_SWIGLU", "1") != "0"
# q | k | v of the fused block as ONE forward GEMM over the stacked dequantised weights (three 25 us launches -> one);
# VPT_FUSE_QKV=0 = three calls
FUSE_QKV = os.environ.get("VPT_FUSE_QKV", "1") != "0"
# input gradient of a dense frozen bf16 linear (patch embed, final layer) as a forward call on the cached transposed weight
# (CTA-pair kernel) instead of the 1-CTA backward kernel
DENSE_BWD_VIA_TRANSPOSE = os.environ.get("VPT_DENSE_BWD_TRANSPOSE", "1") != "0"
# the step's glue as library kernels: noise preparation (noise_mix), the patch-token slice (token_prefix / packed_tokens), the
# upstream-gradient scaling of the loss, the final layer's MLP through the SwiGLU epilogues; off = the ATen ops these
# replaced (A/B measurements, VPT_FUSED_GLUE=0)
FUSED_GLUE = os.environ.get("VPT_FUSED_GLUE", "1") != "0"
_SCRATCH: dict[tuple, torch.Tensor] = {}


def _weight_scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per (device, stream) workspace for the dequantised weight (vpt_linear_scratch_bytes); kernels on one stream run
    in order, so one buffer per stream is enough.  Grown geometrically, never freed."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 16 << 20), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


_ARENA: dict[tuple, torch.Tensor] = {}


def dequant_block(weights: list, downs: list, ups: list, transposed: bool, arena_tag=None) -> list[torch.Tensor] | None:
    """Dequantises the NF4 weights of one transformer block with ONE launch (vpt_nf4_dequant_batch), each into its own
    slot of a per-stream arena; the slots are then handed to linear_raw(scratch=...).  Returns None when the block is not
    made of NF4 weights only (the caller falls back to per-call dequantisation).  `arena_tag` names the arena explicitly
    (default: one per launching stream): the prefetcher alternates two arenas so that the next block's weights can be
    produced while the current block still reads its own."""
    if not weights or len(weights) > 8 or not all(isinstance(w, Nf4Tensors) for w in weights):
        return None
    lib = _lib.load()
    dev = weights[0].packed.device
    # one direction's bytes per slot: forward slots of N % 128 == 0 weights are then gap-free (see fused_rows_view)
    sizes = [(int(lib.vpt_linear_scratch_bytes_dir(w.shape[0], w.shape[1], int(transposed))) + 255) // 256 * 256 for w in weights]
    key = (dev.index, arena_tag if arena_tag is not None else torch.cuda.current_stream(dev).cuda_stream, bool(transposed))
    arena = _ARENA.get(key)
    if arena is None or arena.numel() < sum(sizes):
        arena = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
        _ARENA[key] = arena
    arr = (DequantItemC * len(weights))()
    slots, off = [], 0
    for it, w, d, u, sz in zip(arr, weights, downs, ups, sizes):
        slot = arena[off:off + sz]
        off += sz
        slots.append(slot)
        it.w = w.c_struct()
        it.w_scratch = slot.data_ptr()
        it.scratch_bytes = sz
        if transposed and d is not None:
            it.lora_down, it.ld_lora_down, it.lora_up = d.data_ptr(), d.stride(0), u.data_ptr()
    _lib.call("vpt_nf4_dequant_batch", arr, len(weights), int(transposed), _stream())
    if GEMM_TIMER is not None:
        GEMM_TIMER.append({"kind": "dequant", "weights": list(weights), "downs": list(downs), "ups": list(ups),
                           "transposed": bool(transposed), "slot_ptrs": [s_.data_ptr() for s_ in slots],
                           "arena_tag": arena_tag})
    return slots


def fused_rows_view(slots: list[torch.Tensor], n_rows: int, k: int) -> torch.Tensor | None:
    """The forward slots of several [n_rows, k] weights, laid out back to back by dequant_block, as ONE bf16 [len * n_rows, k]
    weight (row pitch k rounded up to 8) -- what the sectioned forward call takes.  None when they are not adjacent."""
    ldk = (k + 7) // 8 * 8
    need = n_rows * ldk * 2
    for a, b in zip(slots, slots[1:]):
        if a.numel() != need or b.data_ptr() != a.data_ptr() + need:
            return None
    if slots[-1].numel() < need:
        return None
    base = slots[0]
    flat = base.new_empty(0).set_(base.untyped_storage(), base.storage_offset(), (need * len(slots),), (1,))
    return flat.view(torch.bfloat16).view(len(slots) * n_rows, ldk)[:, :k]


def stacked(tensors: list[torch.Tensor]) -> torch.Tensor:
    """Row-wise concatenation of 2-D (or 1-D) tensors; a zero-copy view when they already sit back to back in one storage
    (the q / k / v LoRA matrices inside train.FlatLoRA's buffer), a copy otherwise."""
    first = tensors[0]
    ok = all(t.is_contiguous() and t.dtype == first.dtype and t.shape[1:] == first.shape[1:] for t in tensors)
    if ok:
        for a, b in zip(tensors, tensors[1:]):
            if a.untyped_storage().data_ptr() != b.untyped_storage().data_ptr() or \
                    b.data_ptr() != a.data_ptr() + a.numel() * a.element_size():
                ok = False
                break
    if not ok:
        return torch.cat(tensors, dim=0)
    rows = sum(t.shape[0] for t in tensors)
    shape = (rows,) + tuple(first.shape[1:])
    return first.new_empty(0).set_(first.untyped_storage(), first.storage_offset(), shape, first.stride())


class DequantPrefetcher:
    """Produces the NEXT fused block's dequantised weights on a side stream while the current block computes.

    The batched dequantisation is latency-bound (18 MB in ~16 us, 0.4 ms of a JiT-B step when it sits in line with the
    GEMMs); its inputs are frozen (forward) or fixed for the whole backward pass (the transposed LoRA copies), so block i+1's
    launch can run under block i's first non-persistent kernels.  Two arenas alternate by block parity; the side stream
    first waits for everything already on the main stream (the previous user of that arena has finished by then), the
    consumer waits for the producer's event.  Works under CUDA-graph capture (the waits become graph edges)."""

    def __init__(self):
        self.side: torch.cuda.Stream | None = None
        self.pending: dict = {}

    def issue(self, key, weights, downs, ups, transposed: bool, parity: int) -> None:
        dev = weights[0].packed.device
        main = torch.cuda.current_stream(dev)
        if self.side is None or self.side.device != dev:
            self.side = torch.cuda.Stream(device=dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            slots = dequant_block(weights, downs, ups, transposed, arena_tag=("pf", main.cuda_stream, parity))
            ev = torch.cuda.Event()
            ev.record(self.side)
        if slots is not None:
            self.pending[key] = (slots, ev)

    def take(self, key):
        item = self.pending.pop(key, None)
        if item is None:
            return None
        slots, ev = item
        torch.cuda.current_stream(slots[0].device).wait_event(ev)
        return slots

    def reset(self) -> None:
        """Join and forget whatever is still pending (an abandoned forward / backward pass)."""
        if self.pending and self.side is not None:
            torch.cuda.current_stream(self.side.device).wait_stream(self.side)
        self.pending.clear()


PREFETCH = DequantPrefetcher()
PREFETCH_DEQUANT = os.environ.get("VPT_PREFETCH_DEQUANT", "1") != "0"

# bench.py sets this to a list to bracket every fused-linear launch with CUDA events on the launching stream
# (roofline of the dominant kernel, measured inside real training steps); None = no instrumentation.
GEMM_TIMER: list | None = None

_DT = {torch.bfloat16: _lib.VPT_BF16, torch.float16: _lib.VPT_F16, torch.float32: _lib.VPT_F32}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: torch.Tensor | None) -> C.c_void_p | None:
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vision_pt_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[..., D] -> 2-D [rows, D] view with unit inner stride, 16-byte aligned rows (pitch a multiple of 8 elements).
    Ragged widths (D % 8 != 0, e.g. the 2730-wide SwiGLU hidden of JiT-L) are copied into a padded buffer."""
    t2 = t.reshape(-1, t.shape[-1])
    ok = t2.stride(1) == 1 and (t2.shape[0] == 1 or t2.stride(0) % 8 == 0) and t2.data_ptr() % 16 == 0
    if ok:
        return t2
    rows, d = t2.shape
    if d % 8 == 0:
        return t2.contiguous()
    buf = torch.zeros((rows, (d + 7) // 8 * 8), dtype=t2.dtype, device=t2.device)
    buf[:, :d] = t2
    return buf[:, :d]


def padded_weight(w: torch.Tensor) -> torch.Tensor:
    """[N, K] view of a frozen bf16 weight whose rows start on 16-byte boundaries (pitch = K rounded up to 8): what the
    tensor-map needs when in_features is ragged (3413 = JiT-H's SwiGLU width, 2730 = JiT-L's).  The padded copy is made
    once and kept ON the weight object (so it lives and dies with it -- a cache keyed by address would hand a freed
    weight's copy to whatever is allocated there next) and is rebuilt if the weight is modified in place."""
    N, K = w.shape
    if K % 8 == 0 and w.is_contiguous() and w.data_ptr() % 16 == 0:
        return w
    hit = getattr(w, "_vpt_padded", None)
    if hit is None or hit[0] != w._version or hit[1].device != w.device or hit[2] != w.data_ptr():
        buf = torch.zeros((N, (K + 7) // 8 * 8), dtype=w.dtype, device=w.device)
        buf[:, :K] = w.detach()
        hit = (w._version, buf[:, :K], w.data_ptr())
        w._vpt_padded = hit
    return hit[1]


def transposed_weight(w: torch.Tensor) -> torch.Tensor:
    """[K, N] view (row pitch = N rounded up to 8, zero padding) of the transpose of a frozen bf16 [N, K] weight: with it the
    input gradient dX = dY W of a dense frozen linear is a FORWARD call of the CTA-pair kernel (a K-major operand, as the
    transposed dequantisation gives the NF4 weights) instead of the 1-CTA backward kernel.  Cached on the weight object
    like padded_weight."""
    N, K = w.shape
    hit = getattr(w, "_vpt_transposed", None)
    if hit is None or hit[0] != w._version or hit[1].device != w.device or hit[2] != w.data_ptr():
        buf = torch.zeros((K, (N + 7) // 8 * 8), dtype=w.dtype, device=w.device)
        buf[:, :N] = w.detach().t()
        hit = (w._version, buf[:, :N], w.data_ptr())
        w._vpt_transposed = hit
    return hit[1]


def packed_tokens(x: torch.Tensor) -> torch.Tensor:
    """Contiguous copy of a [B, n, D] slice of a wider token buffer (rows contiguous inside a batch entry, any batch pitch)
    as one 16-byte-vectorised launch; anything else goes through torch."""
    if x.is_contiguous():
        return x
    es = x.element_size()
    if (FUSED_GLUE and x.dim() == 3 and x.is_cuda and x.stride(2) == 1 and x.stride(1) == x.shape[2] and x.stride(0) >= x.shape[1] * x.shape[2]
            and (x.shape[1] * x.shape[2] * es) % 16 == 0 and (x.stride(0) * es) % 16 == 0 and x.data_ptr() % 16 == 0):
        B, n, D = x.shape
        out = torch.empty((B, n, D), dtype=x.dtype, device=x.device)
        _lib.call("vpt_copy_rows", _p(out), n * D * es, _p(x), x.stride(0) * es, B, n * D * es, _stream())
        return out
    return x.contiguous()


class TokenPrefixFn(torch.autograd.Function):
    """x[:, :n] of a [B, L, D] token buffer as a contiguous tensor (JiT.forward: the patch tokens handed to the final layer,
    reference denoiser.py:1115-1124); backward = the gradient in a zero-tailed [B, L, D] buffer.  Two vectorised launches
    where the slice + reshape pair costs torch a strided copy each way plus a fill."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.L = x.shape[1]
        return packed_tokens(x[:, :n])

    @staticmethod
    def backward(ctx, dy):
        B, n, D = dy.shape
        dyc = dy.contiguous()
        dx = torch.empty((B, ctx.L, D), dtype=dy.dtype, device=dy.device)
        es = dy.element_size()
        if dy.is_cuda and (n * D * es) % 16 == 0 and (ctx.L * D * es) % 16 == 0:
            _lib.call("vpt_copy_rows", _p(dx), ctx.L * D * es, _p(dyc), n * D * es, B, n * D * es, _stream())
        else:
            dx[:, :n] = dyc
        copy_token_slots(dx, n, None)
        return dx, None


def token_prefix(x: torch.Tensor, n: int) -> torch.Tensor:
    if n == x.shape[1]:
        return x
    if not FUSED_GLUE:
        return x[:, :n]
    return TokenPrefixFn.apply(x, n)


# ------------------------------------------------------------------------------------------------------- NF4
@dataclass
class Nf4Tensors:
    """Device tensors of one bitsandbytes NF4 weight (see include/vptb200.h: vpt_nf4_weight)."""
    packed: torch.Tensor         # uint8 [(N*K+1)//2, 1]
    absmax: torch.Tensor         # uint8 [N*K/64]
    nested_absmax: torch.Tensor  # fp32
    nested_code: torch.Tensor    # fp32 [256]
    code: torch.Tensor           # fp32 [16]
    offset: float
    shape: tuple[int, int]
    dtype: torch.dtype
    # row-aligned copy for ragged in_features (K % 64 != 0), built lazily on the GPU by vpt_nf4_repack
    packed_rows: torch.Tensor | None = None
    absmax_f32: torch.Tensor | None = None

    @property
    def k_pad(self) -> int:
        return (self.shape[1] + 63) // 64 * 64

    def c_struct(self, for_gemm: bool = False) -> Nf4WeightC:
        rows = stats = None
        kp = 0
        if for_gemm and self.shape[1] % 64 != 0:
            self.ensure_repacked()
            rows, stats, kp = self.packed_rows, self.absmax_f32, self.k_pad
        return Nf4WeightC(_p(self.packed), _p(self.absmax), _p(self.nested_absmax), _p(self.nested_code), _p(self.code),
                          float(self.offset), int(self.shape[0]), int(self.shape[1]), _p(rows), _p(stats), kp)

    def ensure_repacked(self) -> None:
        if self.packed_rows is not None:
            return
        n, k = self.shape
        dev = self.packed.device
        rows = torch.empty((n, self.k_pad // 2), dtype=torch.uint8, device=dev)
        stats = torch.empty((n * k + 63) // 64, dtype=torch.float32, device=dev)
        ws = self.c_struct()
        _lib.call("vpt_nf4_repack", C.byref(ws), _p(rows), _p(stats), self.k_pad, _stream())
        self.packed_rows, self.absmax_f32 = rows, stats

    def to(self, device) -> "Nf4Tensors":
        return Nf4Tensors(self.packed.to(device), self.absmax.to(device), self.nested_absmax.to(device),
                          self.nested_code.to(device), self.code.to(device), self.offset, self.shape, self.dtype)


def nf4_dequantize(w: Nf4Tensors, out_dtype: torch.dtype | None = None) -> torch.Tensor:
    _need_cuda(w.packed)
    dt = out_dtype or w.dtype
    n = w.shape[0] * w.shape[1]
    out = torch.empty(w.shape, dtype=dt, device=w.packed.device)
    ws = w.c_struct()
    _lib.call("vpt_nf4_dequant", C.byref(ws), n, _DT[dt], _p(out), _stream())
    return out


def nf4_quantize(weight: torch.Tensor, nested_code: torch.Tensor, code: torch.Tensor) -> Nf4Tensors:
    """quantize_4bit(weight, quant_type='nf4', blocksize=64, compress_statistics=True) on the GPU."""
    _need_cuda(weight)
    w = weight.detach().contiguous()
    n = w.numel()
    if n % 64 != 0:
        raise ValueError("NF4 quantisation needs a multiple of 64 elements")
    dev = w.device
    nb = n // 64
    packed = torch.empty((n // 2, 1), dtype=torch.uint8, device=dev)
    qabs = torch.empty(nb, dtype=torch.uint8, device=dev)
    nested = torch.empty((nb + 255) // 256, dtype=torch.float32, device=dev)
    offset = torch.empty(1, dtype=torch.float32, device=dev)
    ws = torch.empty(nb, dtype=torch.float32, device=dev)
    ncode = nested_code.to(device=dev, dtype=torch.float32).contiguous()
    _lib.call("vpt_nf4_quantize", _p(w), _DT[w.dtype], n, _p(ncode), _p(packed), _p(qabs), _p(nested), _p(offset), _p(ws),
              _stream())
    return Nf4Tensors(packed, qabs, nested, ncode, code.to(device=dev, dtype=torch.float32).contiguous(),
                      float(offset.item()), (int(weight.shape[0]), int(weight.shape[1])), weight.dtype)


# ------------------------------------------------------------------------------------------------------- linear
def _pad_rank(down: torch.Tensor | None, up: torch.Tensor | None):
    """LoRA matrices in the kernel's shape: rank padded to 16 with zeros, lora_down rows padded to a 16-byte pitch."""
    if down is None:
        return None, None
    r, k = down.shape
    if r > RANK:
        raise NotImplementedError(f"the fused NF4-LoRA kernel handles rank <= {RANK}, got {r}")
    if r == RANK and k % 8 == 0:
        return down.contiguous(), up.contiguous()
    k8 = (k + 7) // 8 * 8
    d = down.new_zeros(RANK, k8)
    d[:r, :k] = down
    if r == RANK:
        return d[:, :k], up.contiguous()
    u = up.new_zeros(up.shape[0], RANK)
    u[:, :r] = up
    return d[:, :k], u


def linear_raw(x2: torch.Tensor, w: Nf4Tensors | torch.Tensor, bias, down, up, scale: float, residual=None,
               want_side: bool = False, backward: bool = False, tile_n: int = 0, reuse_scratch: bool = False,
               scratch: torch.Tensor | None = None, epilogue: int = 0, in2: torch.Tensor | None = None, n_sections: int = 1,
               tape_slot: torch.Tensor | None = None):
    """One call of the fused linear.  forward: x2 [M,K] -> y [M,N]; backward: x2 = dy [M,N] -> dx [M,K].
    `w` is the NF4 tensor set or a plain bf16 [N,K] weight.  Returns (out, side or None) -- or (out, out2, side) with a
    fused SwiGLU epilogue (include/vptb200.h): epilogue=1 (forward of w_2, residual = g) gives (a, u, side), epilogue=2
    (backward of w_3, residual = g, in2 = u) gives (dg, du, side).  n_sections > 1: `w` is a bf16 weight holding several
    linears stacked row-wise (q | k | v over one input), bias / lora_up stacked alike, lora_down 16 rows per section; the side
    tensor then has 16 rows per section.  tape_slot: the dequantisation slot such a stacked weight starts at (bench taping).
    reuse_scratch: the previous call on this stream used the same weight and direction, so the dequantised copy in the
    workspace is still valid and the dequantisation kernel is skipped.  scratch: a slot filled by dequant_block for this
    weight and direction (implies reuse)."""
    _need_cuda(x2)
    if x2.dtype != torch.bfloat16:
        raise TypeError("fused linear runs in bfloat16")
    args = LinearArgsC()
    M = x2.shape[0]
    prefilled = scratch is not None
    if prefilled:
        reuse_scratch = True
    if isinstance(w, Nf4Tensors):
        _need_cuda(w.packed, w.absmax, bias)
        N, K = w.shape
        use_scratch = prefilled or NF4_GEMM_MODE == "scratch" or (NF4_GEMM_MODE == "auto" and M >= NF4_SCRATCH_MIN_M)
        args.w = w.c_struct(for_gemm=not use_scratch)
        args.w_bf16 = None
        if use_scratch:
            if not prefilled:
                scratch = _weight_scratch(x2.device, int(_lib.load().vpt_linear_scratch_bytes(N, K)))
            args.w_scratch = _p(scratch)
            args.ld_scratch = (K + 7) // 8 * 8
            args.scratch_bytes = scratch.numel()
            args.reuse_scratch = int(reuse_scratch)
    else:
        N, K = w.shape
        if w.stride(1) != 1 or w.stride(0) % 8 != 0 or w.stride(0) < K or w.data_ptr() % 16 != 0:
            raise ValueError("bf16 weight: unit inner stride and a row pitch that is a multiple of 8 elements "
                             "(ops.padded_weight makes such a view of a ragged [N, K] weight)")
        args.w = Nf4WeightC(None, None, None, None, None, 0.0, int(N), int(K), None, None, 0)
        args.w_bf16 = _p(w)
        args.ld_scratch = w.stride(0)                # row pitch of the bf16 weight (ABI 4)
    n_out = K if backward else N
    ld_out = (n_out + 7) // 8 * 8
    out_full = torch.empty((M, ld_out), dtype=torch.bfloat16, device=x2.device)
    out = out_full[:, :n_out] if ld_out != n_out else out_full
    # the rank-16 projection, kept for the parameter gradients in the [16, M] layout vpt_lora_grad_batch reads by TMA
    ld_side = (M + 7) // 8 * 8
    side = torch.empty((RANK * max(1, n_sections), ld_side), dtype=torch.bfloat16, device=x2.device) \
        if (want_side and down is not None) else None
    if n_sections > 1:
        if isinstance(w, Nf4Tensors) or backward or epilogue or down is None or down.shape[0] != RANK * n_sections:
            raise ValueError("n_sections: a forward call on a stacked bf16 weight with 16 lora_down rows per section")
        args.n_sections = int(n_sections)
    args.bias = _p(bias)
    args.lora_down = _p(down)
    args.ld_lora_down = down.stride(0) if down is not None else 0
    args.lora_up = _p(up)
    args.scale = float(scale)
    args.inp = _p(x2)
    args.ld_in = x2.stride(0) if M > 1 else x2.shape[1]
    args.out = _p(out_full)
    args.ld_out = ld_out
    args.residual = _p(residual)
    args.ld_res = residual.stride(0) if residual is not None and M > 1 else n_out
    args.side = _p(side)
    args.ld_side = ld_side
    args.M = M
    args.tile_n = tile_n
    out2 = None
    if epilogue:
        if (scratch is None and isinstance(w, Nf4Tensors)) or residual is None or (epilogue == 2 and in2 is None):
            raise ValueError("fused SwiGLU epilogue: needs the scratch (CTA-pair) route, residual = g and, in mode 2, in2 = u")
        out2_full = torch.empty((M, ld_out), dtype=torch.bfloat16, device=x2.device)
        out2 = out2_full[:, :n_out] if ld_out != n_out else out2_full
        args.epilogue = int(epilogue)
        args.out2, args.ld_out2 = _p(out2_full), ld_out
        if in2 is not None:
            args.in2, args.ld_in2 = _p(in2), in2.stride(0) if M > 1 else n_out
    timer = GEMM_TIMER
    if timer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("vpt_nf4lora_linear_bwd_dx" if backward else "vpt_nf4lora_linear_fwd", C.byref(args), _stream())
    if scratch is not None and not reuse_scratch:
        _lib.add_launches(1)          # the per-call dequantisation kernel in front of the GEMM
    if timer is not None:
        e1.record()
        lora = down is not None
        flops = 2.0 * M * K * N + (2.0 * M * RANK * (max(1, n_sections) * K + N) if lora else 0.0)
        if tape_slot is not None:            # a stacked weight made of dequantised NF4 slots counts as the NF4 route it is
            scratch, prefilled = tape_slot, True
        timer.append({"kind": "gemm", "e0": e0, "e1": e1, "flops": flops, "M": M, "K": K, "N": N, "bwd": backward, "lora": lora,
                      "nf4": isinstance(w, Nf4Tensors) or tape_slot is not None, "scratch": scratch is not None,
                      "scratch_ptr": scratch.data_ptr() if prefilled else None, "n_sections": int(n_sections),
                      "call": (x2.shape, x2.stride(0), w, bias, down, up, scale, residual is not None, want_side, backward),
                      "epilogue": int(epilogue)})
    if epilogue:
        return out, out2, side
    return out, side


def lora_grad_batch(items: list[tuple]) -> None:
    """One launch of vpt_lora_grad_batch.  items: (src2 [M,P], [side tensors [16, ld]], [fp32 outputs], transposed);
    out_i += src2^T side_i^T, written as [P,16] or, when transposed, as [16,P]."""
    for lo in range(0, len(items), 16):
        chunk = items[lo:lo + 16]
        arr = (LoraGradItemC * len(chunk))()
        for it, (src2, sides, outs, transposed) in zip(arr, chunk):
            M, P = src2.shape
            it.src = src2.data_ptr()
            it.ld_src = src2.stride(0) if M > 1 else (P + 7) // 8 * 8
            it.M, it.P, it.nsmall = M, P, len(sides)
            for j, (sd, out) in enumerate(zip(sides, outs)):
                if sd.shape[0] != RANK or sd.shape[1] < M or out.dtype != torch.float32 or not out.is_contiguous():
                    raise ValueError("lora_grad_batch: side tensors are [16, >= M] bf16, outputs contiguous fp32")
                it.small_t[j] = sd.data_ptr()
                it.out[j] = out.data_ptr()
            it.ld_small = sides[0].stride(0)
            it.transposed = int(transposed)
            it.ld_out = outs[0].shape[1] if transposed else RANK
        _lib.call("vpt_lora_grad_batch", arr, len(chunk), _stream())


def grad_sink(param: torch.Tensor | None) -> torch.Tensor | None:
    """fp32 slice of the trainer's flat LoRA-gradient buffer for this parameter (vision_pt_b200.train.FlatLoRA), or None.
    When present, the lora_grad kernels accumulate straight into it and autograd receives no gradient for the matrix."""
    return getattr(param, "_vpt_grad32", None) if param is not None else None


def lora_param_grads(dy2, side, x2, dside, down, up, rank: int):
    """lora_down / lora_up gradients of one linear.  Returns (ddown, dup) for autograd, or (None, None) when they were
    accumulated into the flat fp32 buffer."""
    sink_d, sink_u = grad_sink(down), grad_sink(up)
    if sink_d is not None and sink_u is not None and rank == RANK:
        lora_grad_batch([(dy2, [side], [sink_u], False), (x2, [dside], [sink_d], True)])
        return None, None
    N, K = dy2.shape[1], x2.shape[1]
    gup = torch.zeros((N, RANK), dtype=torch.float32, device=dy2.device)
    gdown = torch.zeros((RANK, K), dtype=torch.float32, device=dy2.device)
    lora_grad_batch([(dy2, [side], [gup], False), (x2, [dside], [gdown], True)])
    return gdown[:rank].to(down.dtype), gup[:, :rank].to(up.dtype)


class NF4LoRALinearFn(torch.autograd.Function):
    """y = x W^T + b + (alpha/r) (x A^T) B^T with W in NF4 (or bf16) -- LoRALinear over BnbLinear4bit."""

    @staticmethod
    def forward(ctx, x, w, bias, down, up, scale, residual):
        K = x.shape[-1]
        in_dtype = x.dtype
        xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        x2 = _rows(xb)
        rank = 0 if down is None else down.shape[0]
        dpad, upad = _pad_rank(down, up)
        res2 = _rows(residual) if residual is not None else None
        bias_b = None if bias is None else bias.to(torch.bfloat16)
        y, side = linear_raw(x2, w, bias_b, dpad, upad, scale, res2, want_side=down is not None)
        ctx.w, ctx.scale, ctx.rank, ctx.in_dtype = w, scale, rank, in_dtype
        ctx.lora_params = (down, up)
        ctx.has_res = residual is not None
        ctx.bias_grad = bias is not None and bias.requires_grad
        ctx.save_for_backward(x2, side, dpad, upad)
        N = y.shape[1]
        y = y.reshape(*x.shape[:-1], N)
        return y if in_dtype == torch.bfloat16 else y.to(in_dtype)

    @staticmethod
    def backward(ctx, dy):
        x2, side, dpad, upad = ctx.saved_tensors
        N = dy.shape[-1]
        dy2 = _rows(dy if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16))
        dx = ddown = dup = dbias = None
        lora = dpad is not None
        need_lora_grad = lora and (ctx.needs_input_grad[3] or ctx.needs_input_grad[4])
        dside = None
        if ctx.needs_input_grad[0] and not lora and isinstance(ctx.w, torch.Tensor) and DENSE_BWD_VIA_TRANSPOSE and FUSED_GLUE \
                and not ctx.w.requires_grad:
            dx2, _ = linear_raw(dy2, transposed_weight(ctx.w), None, None, None, 1.0, None)      # dX = dY (W^T)^T
            dx = dx2.reshape(*dy.shape[:-1], x2.shape[1])
            if ctx.in_dtype != torch.bfloat16:
                dx = dx.to(ctx.in_dtype)
        elif ctx.needs_input_grad[0] or need_lora_grad:
            dx2, dside = linear_raw(dy2, ctx.w, None, dpad, upad, ctx.scale, None, want_side=lora, backward=True)
            if ctx.needs_input_grad[0]:
                dx = dx2.reshape(*dy.shape[:-1], x2.shape[1])
                if ctx.in_dtype != torch.bfloat16:
                    dx = dx.to(ctx.in_dtype)
        if need_lora_grad:
            ddown, dup = lora_param_grads(dy2, side, x2, dside, ctx.lora_params[0], ctx.lora_params[1], ctx.rank)
        if ctx.bias_grad:
            dbias = dy2.float().sum(0).to(dy.dtype)
        dres = dy if ctx.has_res else None
        return dx, None, dbias, ddown, dup, None, dres


def nf4_lora_linear(x, w, bias=None, lora_down=None, lora_up=None, scale: float = 1.0, residual=None):
    return NF4LoRALinearFn.apply(x, w, bias, lora_down, lora_up, scale, residual)


# ------------------------------------------------------------------------------------------------------- attention
HEAD_DIMS = (64, 80, 32, 96, 128)     # 64 / 80 = tcgen05 kernels (JiT-B/L, SDXL / JiT-H); the rest run attention_simple.cuh
QKNORM_HEAD_DIMS = (64, 80, 96, 128)


def _at(t: torch.Tensor) -> AttnTensorC:
    # logical layout (B, H, L, 64); any strides with a contiguous last dimension
    return AttnTensorC(_p(t), t.stride(0), t.stride(2), t.stride(1))


def _attn_ok(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError("attention kernels run in bfloat16")
    if t.shape[-1] not in HEAD_DIMS:
        raise NotImplementedError(f"attention kernels exist for head_dim {HEAD_DIMS} (64: tcgen05; the others: CUDA cores)")
    if t.stride(3) != 1 or any(s % 8 for s in t.stride()[:3]) or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def attn_fwd_raw(q, k, v, seqlens_k, scale: float):
    """q,k,v: [B,H,L,hd] views.  Returns (o [B,H,Lq,hd] view over token-major memory, lse2 [B,H,Lq rounded up to 128])."""
    B, H, Lq, hd = q.shape
    Lk = k.shape[2]
    o = torch.empty((B, Lq, H, hd), dtype=torch.bfloat16, device=q.device).permute(0, 2, 1, 3)
    lse2 = torch.empty((B, H, (Lq + 127) // 128 * 128), dtype=torch.float32, device=q.device)
    tq, tk, tv, to = _at(q), _at(k), _at(v), _at(o)
    _lib.call("vpt_attn_fwd", C.byref(tq), C.byref(tk), C.byref(tv), C.byref(to), B, H, Lq, Lk, hd, _p(seqlens_k),
              float(scale), _p(lse2), _stream())
    return o, lse2


def attn_bwd_raw(q, k, v, o, d_o, lse2, seqlens_k, scale: float):
    """Returns (dq fp32, dk bf16, dv bf16), each a [B,H,L,hd] view over token-major memory."""
    B, H, Lq, hd = q.shape
    Lk = k.shape[2]
    dev = q.device
    dq = torch.zeros((B, Lq, H, hd), dtype=torch.float32, device=dev).permute(0, 2, 1, 3)
    dk = torch.empty((B, Lk, H, hd), dtype=torch.bfloat16, device=dev).permute(0, 2, 1, 3)
    dv = torch.empty((B, Lk, H, hd), dtype=torch.bfloat16, device=dev).permute(0, 2, 1, 3)
    delta = torch.empty((B, H, (Lq + 127) // 128 * 128), dtype=torch.float32, device=dev)
    ts = [_at(t) for t in (q, k, v, o, d_o, dq, dk, dv)]
    _lib.call("vpt_attn_bwd", *[C.byref(t) for t in ts], B, H, Lq, Lk, hd, _p(seqlens_k), float(scale), _p(lse2), _p(delta),
              _stream())
    if hd not in (64, 80):
        _lib.add_launches(2)          # delta + dQ + dK + dV kernels on the CUDA-core route
    return dq, dk, dv


class AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, seqlens_k, scale):
        q, k, v = _attn_ok(q), _attn_ok(k), _attn_ok(v)
        o, lse2 = attn_fwd_raw(q, k, v, seqlens_k, scale)
        ctx.save_for_backward(q, k, v, o, lse2, seqlens_k)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse2, seqlens_k = ctx.saved_tensors
        dq, dk, dv = attn_bwd_raw(q, k, v, o, _attn_ok(d_o), lse2, seqlens_k, ctx.scale)
        return dq.to(torch.bfloat16), dk, dv, None, None


def attention(q, k, v, seqlens_k=None, scale: float | None = None):
    """softmax(q k^T * scale + key-padding mask) v for [B,H,L,64] bf16 tensors; seqlens_k int32 [B] or None."""
    _need_cuda(q, k, v)
    if scale is None:
        scale = q.shape[-1] ** -0.5
    return AttentionFn.apply(q, k, v, seqlens_k, float(scale))


# ------------------------------------------------------------------------------------------------------- norms
def rmsnorm_fwd_raw(x2, w, eps: float, want_rstd: bool = True):
    rows, D = x2.shape
    y = torch.empty((rows, D), dtype=torch.bfloat16, device=x2.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x2.device) if want_rstd else None
    _lib.call("vpt_rmsnorm_fwd", _p(x2), _p(w), _p(y), _p(rstd), rows, D, x2.stride(0) if rows > 1 else D, D, float(eps),
              _stream())
    return y, rstd


def rmsnorm_bwd_raw(dy2, x2, w, rstd, dres2, eps: float, dw=None):
    rows, D = x2.shape
    assert dy2.is_contiguous() and x2.is_contiguous() and (dres2 is None or dres2.is_contiguous())
    dx = torch.empty((rows, D), dtype=torch.bfloat16, device=x2.device)
    _lib.call("vpt_rmsnorm_bwd", _p(dy2), _p(x2), _p(w), _p(rstd), _p(dres2), _p(dx), _p(dw), rows, D, D, float(eps),
              _stream())
    return dx


class RMSNormFn(torch.autograd.Function):
    """FP32RMSNorm.forward: bf16(rms_norm(x.float()) * w)."""

    @staticmethod
    def forward(ctx, x, w, eps):
        if x.dtype != torch.bfloat16:
            raise TypeError("fused RMSNorm runs on bfloat16 activations")
        D = x.shape[-1]
        if D == 64 and x.dim() == 4:
            x2 = x.contiguous().reshape(-1, D)
        else:
            x2 = _rows(x)
        wb = None if w is None else w.to(torch.bfloat16)
        y, rstd = rmsnorm_fwd_raw(x2, wb, eps)
        ctx.save_for_backward(x2.contiguous(), wb, rstd)
        ctx.eps = eps
        ctx.w_grad = w is not None and w.requires_grad
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, wb, rstd = ctx.saved_tensors
        dy2 = dy.to(torch.bfloat16).reshape(x2.shape).contiguous()
        dw = torch.zeros(x2.shape[1], dtype=torch.float32, device=x2.device) if ctx.w_grad else None
        dx = rmsnorm_bwd_raw(dy2, x2, wb, rstd, None, ctx.eps, dw)
        return dx.reshape(dy.shape), (dw.to(wb.dtype) if dw is not None else None), None


def rms_norm(x, weight, eps: float = 1e-6):
    _need_cuda(x)
    return RMSNormFn.apply(x, weight, eps)


def qknorm_rope_fwd_raw(x2, w, cos_sin, H: int, L: int, eps: float):
    tokens = x2.shape[0]
    hd = x2.shape[1] // H
    y = torch.empty((tokens, H * hd), dtype=torch.bfloat16, device=x2.device)
    _lib.call("vpt_qknorm_rope_fwd", _p(x2), _p(w), _p(cos_sin), _p(y), tokens, H, L, hd,
              x2.stride(0) if tokens > 1 else H * hd, H * hd, float(eps), _stream())
    return y


def qknorm_rope_bwd_raw(dy2, x2, w, cos_sin, H: int, L: int, eps: float, dw=None):
    tokens = x2.shape[0]
    hd = x2.shape[1] // H
    dx = torch.empty((tokens, H * hd), dtype=torch.bfloat16, device=x2.device)
    _lib.call("vpt_qknorm_rope_bwd", _p(dy2), int(dy2.dtype == torch.float32), _p(x2), _p(w), _p(cos_sin), _p(dx), _p(dw),
              tokens, H, L, hd, dy2.stride(0) if tokens > 1 else H * hd, x2.stride(0) if tokens > 1 else H * hd, H * hd,
              float(eps), _stream())
    return dx


class QKNormRopeFn(torch.autograd.Function):
    """q_norm (RMS over head_dim) followed by apply_rope, on token-major [B, L, H, 64]."""

    @staticmethod
    def forward(ctx, x, w, cos_sin, eps):
        B, L, H, hd = x.shape
        x2 = _rows(x.reshape(B * L, H * hd))
        wb = w.to(torch.bfloat16)
        y = qknorm_rope_fwd_raw(x2, wb, cos_sin, H, L, eps)
        ctx.save_for_backward(x2, wb, cos_sin)
        ctx.dims = (B, L, H, hd, eps, w.requires_grad)
        return y.reshape(B, L, H, hd)

    @staticmethod
    def backward(ctx, dy):
        x2, wb, cos_sin = ctx.saved_tensors
        B, L, H, hd, eps, wg = ctx.dims
        dy2 = dy.reshape(B * L, H * hd)
        if dy2.dtype not in (torch.float32, torch.bfloat16):
            dy2 = dy2.to(torch.bfloat16)
        dy2 = dy2.contiguous()
        dw = torch.zeros(hd, dtype=torch.float32, device=x2.device) if wg else None
        dx = qknorm_rope_bwd_raw(dy2, x2, wb, cos_sin, H, L, eps, dw)
        return dx.reshape(B, L, H, hd), (dw.to(wb.dtype) if dw is not None else None), None, None


def qknorm_rope(x, weight, cos_sin, eps: float = 1e-6):
    return QKNormRopeFn.apply(x, weight, cos_sin, eps)


# ------------------------------------------------------------------------------------------------------- SwiGLU
def swiglu_fwd_raw(g2, u2):
    """g2, u2: [rows, F] views whose row pitch covers F rounded up to 8 (see _rows / linear_raw)."""
    rows, F = g2.shape
    ld = (F + 7) // 8 * 8
    a_full = torch.empty((rows, ld), dtype=torch.bfloat16, device=g2.device)
    _lib.call("vpt_swiglu_fwd", _p(g2), _p(u2), _p(a_full), rows, F, g2.stride(0), u2.stride(0), ld, _stream())
    return a_full[:, :F] if ld != F else a_full


def swiglu_bwd_raw(da2, g2, u2):
    rows, F = g2.shape
    ld = (F + 7) // 8 * 8
    dg_full = torch.empty((rows, ld), dtype=torch.bfloat16, device=g2.device)
    du_full = torch.empty((rows, ld), dtype=torch.bfloat16, device=g2.device)
    _lib.call("vpt_swiglu_bwd", _p(da2), _p(g2), _p(u2), _p(dg_full), _p(du_full), rows, F, da2.stride(0), g2.stride(0),
              u2.stride(0), ld, ld, _stream())
    if ld != F:
        return dg_full[:, :F], du_full[:, :F]
    return dg_full, du_full


class SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, u):
        g2, u2 = _rows(g), _rows(u)
        a = swiglu_fwd_raw(g2, u2)
        ctx.save_for_backward(g2, u2)
        return a.reshape(g.shape)

    @staticmethod
    def backward(ctx, da):
        g2, u2 = ctx.saved_tensors
        dg, du = swiglu_bwd_raw(_rows(da), g2, u2)
        return dg.reshape(da.shape), du.reshape(da.shape)


def swiglu(g, u):
    return SwiGLUFn.apply(g, u)


class DenseSwiGLUFn(torch.autograd.Function):
    """w_3(silu(w_1 x) * (w_2 x)) for three FROZEN bf16 linears (the final layer's MLP, reference jit/denoiser.py:498-506,
    535-543) with the gate in the epilogues of the CTA-pair GEMMs, like the fused block: forward = w_1, w_2 (+ SwiGLU
    forward), w_3; backward = w_3^T (+ SwiGLU backward), w_1^T, w_2^T (+ the sum of the two input gradients as its residual)
    -- six launches, no stand-alone gate / add kernels.  Weights come as padded_weight views, the backward uses their cached
    transposes."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w3, b3):
        x2 = _rows(x)
        g, _ = linear_raw(x2, w1, b1, None, None, 1.0)
        a, u, _ = linear_raw(x2, w2, b2, None, None, 1.0, g, epilogue=1)
        y, _ = linear_raw(a, w3, b3, None, None, 1.0)
        ctx.weights = (w1, w2, w3)
        ctx.save_for_backward(g, u)
        return y.reshape(*x.shape[:-1], y.shape[1])

    @staticmethod
    def backward(ctx, dy):
        g, u = ctx.saved_tensors
        w1, w2, w3 = ctx.weights
        dy2 = _rows(dy)
        dg, du, _ = linear_raw(dy2, transposed_weight(w3), None, None, None, 1.0, g, epilogue=2, in2=u)
        dx, _ = linear_raw(dg, transposed_weight(w1), None, None, None, 1.0)
        dx, _ = linear_raw(du, transposed_weight(w2), None, None, None, 1.0, dx)
        return dx.reshape(*dy.shape[:-1], dx.shape[1]), None, None, None, None, None, None


def dense_swiglu(x, w1, b1, w2, b2, w3, b3):
    """None when the three linears do not qualify (trainable, not bf16 on a CUDA device): the caller composes the ops."""
    ws, bs = (w1, w2, w3), (b1, b2, b3)
    if not (FUSE_SWIGLU and FUSED_GLUE and x.is_cuda and x.dtype == torch.bfloat16
            and all(w.dtype == torch.bfloat16 and not w.requires_grad for w in ws)
            and all(b is None or (b.dtype == torch.bfloat16 and not b.requires_grad) for b in bs)):
        return None
    return DenseSwiGLUFn.apply(x, padded_weight(w1), b1, padded_weight(w2), b2, padded_weight(w3), b3)


# ------------------------------------------------------------------------------------------------------- adaLN
class LNModulateFn(torch.autograd.Function):
    """FP32LayerNorm(no affine)(x) * (1 + scale[:, None]) + shift[:, None]   x [B,L,D], scale/shift [B,D]."""

    @staticmethod
    def forward(ctx, x, scale, shift, eps):
        B, L, D = x.shape
        xc, sc, sh = x.contiguous(), scale.contiguous(), shift.contiguous()
        y = torch.empty_like(xc)
        mean = torch.empty(B * L, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B * L, dtype=torch.float32, device=x.device)
        _lib.call("vpt_ln_modulate_fwd", _p(xc), _p(sc), _p(sh), _p(y), _p(mean), _p(rstd), B * L, L, D, float(eps), _stream())
        ctx.save_for_backward(xc, sc, mean, rstd)
        ctx.mod_grad = scale.requires_grad or shift.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, sc, mean, rstd = ctx.saved_tensors
        B, L, D = xc.shape
        dyc = dy.contiguous()
        dx = torch.empty_like(xc)
        dsc = dsh = None
        if ctx.mod_grad:
            dsc = torch.zeros((B, D), dtype=torch.float32, device=xc.device)
            dsh = torch.zeros((B, D), dtype=torch.float32, device=xc.device)
        _lib.call("vpt_ln_modulate_bwd", _p(dyc), _p(xc), _p(sc), _p(mean), _p(rstd), _p(dx), _p(dsc), _p(dsh), B * L, L, D,
                  _stream())
        return dx, (dsc.to(sc.dtype) if dsc is not None else None), (dsh.to(sc.dtype) if dsh is not None else None), None


def ln_modulate(x, scale, shift, eps: float = 1e-5):
    _need_cuda(x)
    return LNModulateFn.apply(x, scale, shift, eps)


class GateResidualFn(torch.autograd.Function):
    """x + h * gate[:, None]   x,h [B,L,D], gate [B,D]."""

    @staticmethod
    def forward(ctx, x, h, gate):
        B, L, D = x.shape
        xc, hc, gc = x.contiguous(), h.contiguous(), gate.contiguous()
        y = torch.empty_like(xc)
        _lib.call("vpt_gate_residual_fwd", _p(xc), _p(hc), _p(gc), _p(y), B * L, L, D, _stream())
        ctx.save_for_backward(hc, gc)
        ctx.gate_grad = gate.requires_grad
        return y

    @staticmethod
    def backward(ctx, dy):
        hc, gc = ctx.saved_tensors
        B, L, D = hc.shape
        dyc = dy.contiguous()
        dh = torch.empty_like(hc)
        dg = torch.zeros((B, D), dtype=torch.float32, device=hc.device) if ctx.gate_grad else None
        _lib.call("vpt_gate_residual_bwd", _p(dyc), _p(hc), _p(gc), _p(dh), _p(dg), B * L, L, D, _stream())
        return dy, dh, (dg.to(gc.dtype) if dg is not None else None)


def gate_residual(x, h, gate):
    _need_cuda(x)
    return GateResidualFn.apply(x, h, gate)


# ------------------------------------------------------------------------------------------------------- patchify
def _patch_call(name, src, dst, B, Cc, H, W, p, order):
    if src.element_size() != 2:
        raise TypeError("patchify kernels move 2-byte elements (bf16 / fp16)")
    _lib.call(name, _p(src), _p(dst), B, Cc, H, W, p, order, _stream())


class PatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, p, order):
        B, Cc, H, W = image.shape
        img = image.contiguous()
        out = torch.empty((B, (H // p) * (W // p), Cc * p * p), dtype=image.dtype, device=image.device)
        _patch_call("vpt_patchify", img, out, B, Cc, H, W, p, order)
        ctx.meta = (B, Cc, H, W, p, order)
        return out

    @staticmethod
    def backward(ctx, d):
        B, Cc, H, W, p, order = ctx.meta
        dc = d.contiguous()
        out = torch.empty((B, Cc, H, W), dtype=d.dtype, device=d.device)
        _patch_call("vpt_unpatchify", dc, out, B, Cc, H, W, p, order)
        return out, None, None


class UnpatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patches, Cc, H, W, p, order):
        B = patches.shape[0]
        pc = patches.contiguous()
        out = torch.empty((B, Cc, H, W), dtype=patches.dtype, device=patches.device)
        _patch_call("vpt_unpatchify", pc, out, B, Cc, H, W, p, order)
        ctx.meta = (B, Cc, H, W, p, order)
        return out

    @staticmethod
    def backward(ctx, d):
        B, Cc, H, W, p, order = ctx.meta
        dc = d.contiguous()
        out = torch.empty((B, (H // p) * (W // p), Cc * p * p), dtype=d.dtype, device=d.device)
        _patch_call("vpt_patchify", dc, out, B, Cc, H, W, p, order)
        return out, None, None, None, None, None


def patchify_op(image, p: int, order: int = 0):
    _need_cuda(image)
    return PatchifyFn.apply(image, p, order)


def unpatchify_op(patches, channels: int, height: int, width: int, p: int, order: int = 0):
    _need_cuda(patches)
    return UnpatchifyFn.apply(patches, channels, height, width, p, order)


# ------------------------------------------------------------------------------------------------------- loss / optimiser
class FlowLossFn(torch.autograd.Function):
    """treat_loss for model_pred == "image": MSE on images (mode 0) or on velocities (mode 1); one kernel computes the
    loss and d loss / d pred."""

    @staticmethod
    def forward(ctx, pred, clean, noisy, timestep, mode, clamp_eps):
        if pred.dtype != torch.bfloat16:
            raise TypeError("flow_loss takes the bf16 model prediction")
        pc = pred.contiguous()
        cc = clean.contiguous()
        nc = noisy.contiguous().to(cc.dtype) if noisy is not None else None
        B = pc.shape[0]
        per = pc.numel() // B
        loss = torch.zeros(1, dtype=torch.float32, device=pc.device)
        dpred = torch.empty_like(pc) if pred.requires_grad else None
        t32 = timestep.float().contiguous() if timestep is not None else None
        _lib.call("vpt_flow_loss", _p(pc), _p(cc), _p(nc), _DT[cc.dtype], _p(t32), B, per, int(mode), float(clamp_eps),
                  _p(loss), _p(dpred), _stream())
        ctx.save_for_backward(dpred)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        (dpred,) = ctx.saved_tensors
        if dloss.numel() != 1 or not FUSED_GLUE:
            return dpred * dloss.to(dpred.dtype), None, None, None, None, None
        out = torch.empty_like(dpred)          # = dpred * dloss.to(bf16), without the strided broadcast kernel
        _lib.call("vpt_scale_by_scalar", _p(dpred), _p(dloss.detach().float().reshape(1)), _p(out), dpred.numel(), _stream())
        return out, None, None, None, None, None


def flow_loss(pred, clean, noisy=None, timestep=None, loss_target: str = "image", clamp_eps: float = 0.05):
    _need_cuda(pred, clean)
    return FlowLossFn.apply(pred, clean, noisy, timestep, 1 if loss_target == "velocity" else 0, clamp_eps)


def noise_mix(latents: torch.Tensor, randn: torch.Tensor, timestep: torch.Tensor, noise_scale: float = 1.0,
              clean_at_zero: bool = False, want_bf16: bool = True):
    """prepare_scaled_noised_latents (reference src/modules/loss/flow_match.py:60-74) for a given `randn =
    torch.randn_like(latents)`, one kernel instead of five elementwise passes, bit-identical to the op-by-op form
    (`timestep` is rounded to the latents' dtype first, as a caller passing `timestep.to(latents.dtype)` would).
    Returns (noisy in the latents' dtype, its bf16 copy for the denoiser -- the same tensor when the latents are bf16)."""
    _need_cuda(latents, randn, timestep)
    if latents.dtype not in _DT or randn.dtype != latents.dtype or randn.shape != latents.shape:
        raise TypeError("noise_mix: latents / randn of one shape and dtype (bf16, fp16 or fp32)")
    x, z = latents.contiguous(), randn.contiguous()
    B = x.shape[0]
    if timestep.numel() != B:
        raise ValueError("noise_mix: one timestep per sample")
    t32 = timestep.reshape(B).float().contiguous()
    noisy = torch.empty_like(x)
    as_bf16 = torch.empty_like(x, dtype=torch.bfloat16) if want_bf16 and x.dtype != torch.bfloat16 else None
    _lib.call("vpt_noise_mix", _p(x), _p(z), _DT[x.dtype], _p(t32), B, x.numel() // B, float(noise_scale), int(clean_at_zero),
              _p(noisy), _p(as_bf16), _stream())
    return noisy, (as_bf16 if as_bf16 is not None else noisy)


def grad_sumsq(grad32: torch.Tensor, scale: float, out: torch.Tensor) -> None:
    _lib.call("vpt_grad_sumsq", _p(grad32), grad32.numel(), float(scale), _p(out), _stream())


def adamw_step(param, grad32, exp_avg, exp_avg_sq, step_t, lr, betas, eps, weight_decay, grad_scale=1.0, sumsq=None,
               max_norm=0.0, zero_grad=True) -> None:
    if param.dtype != torch.bfloat16 or grad32.dtype != torch.float32:
        raise TypeError("adamw_step: bf16 parameters with fp32 gradients")
    _lib.call("vpt_adamw_step", _p(param), _p(grad32), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr), float(betas[0]),
              float(betas[1]), float(eps), float(weight_decay), float(grad_scale), _p(sumsq), float(max_norm), _p(step_t),
              int(zero_grad), _stream())


def radam_schedulefree_step(param, grad32, z, exp_avg_sq, sched, coef, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                            r=0.0, weight_lr_power=2.0, silent_sgd_phase=True, grad_scale=1.0, sumsq=None, max_norm=0.0,
                            zero_grad=True) -> None:
    """schedulefree.RAdamScheduleFree.step over the flat buffers (include/vptb200.h).  `sched` float64[4] zeros at the
    start of a run (steps done, lr_max, weight_sum, scheduled_lr), `coef` float32[8] scratch, `z` fp32 clone of `param`."""
    if param.dtype != torch.bfloat16 or grad32.dtype != torch.float32 or z.dtype != torch.float32 or sched.dtype != torch.float64:
        raise TypeError("radam_schedulefree_step: bf16 parameters, fp32 gradients / z, float64 schedule state")
    _lib.call("vpt_radam_schedulefree_step", _p(param), _p(grad32), _p(z), _p(exp_avg_sq), param.numel(), float(lr),
              float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(r), float(weight_lr_power),
              int(silent_sgd_phase), float(grad_scale), _p(sumsq), float(max_norm), _p(sched), _p(coef), int(zero_grad),
              _stream())


def radam_schedulefree_swap(param, z, beta1: float, to_eval: bool) -> None:
    """optimizer.eval() / optimizer.train() of the package: y -> x (the averaged iterate) and back."""
    _lib.call("vpt_radam_schedulefree_swap", _p(param), _p(z), param.numel(), float(beta1), int(to_eval), _stream())


def copy_token_slots(dst3: torch.Tensor, start: int, src3: torch.Tensor | None) -> None:
    """dst3[:, start:] = src3 (or 0 when src3 is None) for a [B, L, D] token buffer: the per-block context-slot refresh of
    JiT (reference denoiser.py:1092-1113) as one 16-byte-vectorised launch; falls back to torch for odd alignments."""
    B, L, D = dst3.shape
    n = L - start
    if n <= 0:
        return
    view = dst3[:, start:]
    es = dst3.element_size()
    ok = (dst3.is_cuda and dst3.stride(2) == 1 and dst3.stride(1) == D and (n * D * es) % 16 == 0
          and (dst3.stride(0) * es) % 16 == 0 and view.data_ptr() % 16 == 0)
    if src3 is not None:
        ok = ok and src3.dtype == dst3.dtype and tuple(src3.shape) == (B, n, D) and src3.stride(2) == 1 and src3.stride(1) == D \
            and (src3.stride(0) * es) % 16 == 0 and src3.data_ptr() % 16 == 0
    if not ok:
        if src3 is None:
            view.zero_()
        else:
            view.copy_(src3)
        return
    _lib.call("vpt_copy_rows", _p(view), dst3.stride(0) * es, _p(src3), 0 if src3 is None else src3.stride(0) * es, B,
              n * D * es, _stream())


# ------------------------------------------------------------------------------------------------------- other block families
# Kernels of csrc/blocks_ext.cuh: what the SDXL / CogView4 / JiT-extension blocks need beyond the JiT block's own kernels.
ACT_KINDS = {"silu": 0, "gelu": 1, "gelu_tanh": 2, "gelu_pytorch_tanh": 2}


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm / FP32LayerNorm with (or without) affine parameters over the last dimension of a bf16 tensor."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        if x.dtype != torch.bfloat16:
            raise TypeError("fused LayerNorm runs on bfloat16 activations")
        D = x.shape[-1]
        x2 = x.reshape(-1, D).contiguous()
        rows = x2.shape[0]
        wb = None if w is None else w.detach().to(torch.bfloat16).contiguous()
        bb = None if b is None else b.detach().to(torch.bfloat16).contiguous()
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        _lib.call("vpt_layernorm_fwd", _p(x2), _p(wb), _p(bb), _p(y), _p(mean), _p(rstd), rows, D, float(eps), _stream())
        ctx.save_for_backward(x2, wb, mean, rstd)
        ctx.grads = (w is not None and w.requires_grad, b is not None and b.requires_grad, None if w is None else w.dtype)
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, wb, mean, rstd = ctx.saved_tensors
        rows, D = x2.shape
        wg, bg, wdt = ctx.grads
        dy2 = dy.to(torch.bfloat16).reshape(rows, D).contiguous()
        dx = torch.empty_like(x2)
        both = D > 2048 and (wg or bg)          # the wide-row kernel accumulates dw and db together
        dw = torch.zeros(D, dtype=torch.float32, device=x2.device) if (wg or both) else None
        db = torch.zeros(D, dtype=torch.float32, device=x2.device) if (bg or both) else None
        _lib.call("vpt_layernorm_bwd", _p(dy2), _p(x2), _p(wb), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db), rows, D, _stream())
        return dx.reshape(dy.shape), (dw.to(wdt) if wg else None), (db.to(wdt) if bg else None), None


def layer_norm(x, weight=None, bias=None, eps: float = 1e-5):
    _need_cuda(x)
    return LayerNormFn.apply(x, weight, bias, eps)


class GatedActFn(torch.autograd.Function):
    """bf16(bf16(act(gate)) * h): GeGLU (h, gate = halves of one projection: views with the projection's row pitch are fine)."""

    @staticmethod
    def forward(ctx, h, gate, kind):
        F_ = h.shape[-1]
        h2, g2 = _rows(h), _rows(gate)
        rows = h2.shape[0]
        ld = (F_ + 7) // 8 * 8
        a_full = torch.empty((rows, ld), dtype=torch.bfloat16, device=h.device)
        _lib.call("vpt_gated_act_fwd", _p(h2), _p(g2), _p(a_full), rows, F_, h2.stride(0) if rows > 1 else ld,
                  g2.stride(0) if rows > 1 else ld, ld, int(kind), _stream())
        ctx.save_for_backward(h2, g2)
        ctx.kind = int(kind)
        a = a_full[:, :F_] if ld != F_ else a_full
        return a.reshape(h.shape)

    @staticmethod
    def backward(ctx, da):
        h2, g2 = ctx.saved_tensors
        rows, F_ = h2.shape
        ld = (F_ + 7) // 8 * 8
        da2 = _rows(da.to(torch.bfloat16))
        dh = torch.empty((rows, ld), dtype=torch.bfloat16, device=h2.device)
        dg = torch.empty((rows, ld), dtype=torch.bfloat16, device=h2.device)
        one = lambda t: t.stride(0) if rows > 1 else ld
        _lib.call("vpt_gated_act_bwd", _p(da2), _p(h2), _p(g2), _p(dh), _p(dg), rows, F_, one(da2), one(h2), one(g2), ld, ld,
                  ctx.kind, _stream())
        if ld != F_:
            dh, dg = dh[:, :F_], dg[:, :F_]
        return dh.reshape(da.shape), dg.reshape(da.shape), None


def gated_act(h, gate, kind: str = "gelu"):
    _need_cuda(h, gate)
    return GatedActFn.apply(h, gate, ACT_KINDS[kind])


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind):
        xc = x.contiguous()
        if xc.numel() % 8 != 0 or xc.dtype != torch.bfloat16:
            raise ValueError("fused activation: bf16 tensors with a multiple of 8 elements")
        y = torch.empty_like(xc)
        _lib.call("vpt_act_fwd", _p(xc), _p(y), xc.numel(), int(kind), _stream())
        ctx.save_for_backward(xc)
        ctx.kind = int(kind)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dyc = dy.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(xc)
        _lib.call("vpt_act_bwd", _p(dyc), _p(xc), _p(dx), xc.numel(), ctx.kind, _stream())
        return dx, None


def activation(x, kind: str = "gelu_tanh"):
    _need_cuda(x)
    return ActFn.apply(x, ACT_KINDS[kind])


def _rope_half_call(x4, cos, sin, l0, inverse):
    B, L, H, hd = x4.shape
    xc = x4.contiguous()
    y = torch.empty_like(xc)
    _lib.call("vpt_rope_half", _p(xc), _p(cos), _p(sin), _p(y), B * L, L, H, hd, int(l0), H * hd, H * hd, int(inverse), _stream())
    return y


class RopeHalfFn(torch.autograd.Function):
    """CogView4 apply_rotary_emb on token-major [B, L, H, hd]; tokens before `l0` (the text stream) pass through."""

    @staticmethod
    def forward(ctx, x4, cos, sin, l0):
        ctx.save_for_backward(cos, sin)
        ctx.l0 = l0
        return _rope_half_call(x4, cos, sin, l0, False)

    @staticmethod
    def backward(ctx, dy):
        cos, sin = ctx.saved_tensors
        return _rope_half_call(dy.to(torch.bfloat16), cos, sin, ctx.l0, True), None, None, None


def rope_half(x4, cos, sin, l0: int = 0):
    """x4 [B, L, H, hd] bf16; cos / sin fp32 [L - l0, hd] (contiguous)."""
    _need_cuda(x4)
    if x4.dtype != torch.bfloat16 or x4.shape[-1] % 16 != 0:
        raise ValueError("rope_half: bf16 [B, L, H, hd] with hd a multiple of 16")
    cos = cos.to(device=x4.device, dtype=torch.float32).contiguous()
    sin = sin.to(device=x4.device, dtype=torch.float32).contiguous()
    if cos.shape != (x4.shape[1] - l0, x4.shape[3]) or sin.shape != cos.shape:
        raise ValueError("rope_half: cos / sin must be [L - l0, head_dim]")
    return RopeHalfFn.apply(x4, cos, sin, int(l0))


class PopeFn(torch.autograd.Function):
    """apply_pope on token-major [B, L, H, d] -> [B, L, H, 2d] (interleaved re / im); frozen learned bias only."""

    @staticmethod
    def forward(ctx, x4, cos_sin, bias):
        B, L, H, d = x4.shape
        xc = x4.contiguous()
        y = torch.empty((B, L, H, 2 * d), dtype=torch.bfloat16, device=x4.device)
        _lib.call("vpt_pope_fwd", _p(xc), _p(cos_sin), _p(bias), _p(y), B * L, L, H, d, H * d, 2 * H * d, _stream())
        ctx.save_for_backward(xc, cos_sin, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, cos_sin, bias = ctx.saved_tensors
        B, L, H, d = xc.shape
        dyc = dy.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(xc)
        _lib.call("vpt_pope_bwd", _p(dyc), _p(xc), _p(cos_sin), _p(bias), _p(dx), B * L, L, H, d, 2 * H * d, H * d, H * d, _stream())
        return dx, None, None


def pope(x4, cos_sin, learned_bias=None):
    """x4 [B, L, H, d] bf16; cos_sin fp32 [L, d, 2]; learned_bias fp32 [H, d] or None (must not require grad here)."""
    _need_cuda(x4)
    if learned_bias is not None and learned_bias.requires_grad:
        raise NotImplementedError("PoPE's learned phase bias is frozen on the fused path")
    if x4.dtype != torch.bfloat16 or x4.shape[-1] % 8 != 0:
        raise ValueError("pope: bf16 [B, L, H, d] with d a multiple of 8")
    cs = cos_sin.to(device=x4.device, dtype=torch.float32).contiguous()
    if cs.shape != (x4.shape[1], x4.shape[3], 2):
        raise ValueError("pope: cos_sin must be [L, d, 2]")
    b = None if learned_bias is None else learned_bias.detach().to(device=x4.device, dtype=torch.float32).contiguous()
    return PopeFn.apply(x4, cs, b)


def _token_gather_call(src3, idx, n_full, scatter, dst=None):
    B, _, D = src3.shape
    sc = src3.contiguous()
    n = idx.numel()
    if dst is None:
        if scatter:
            dst = torch.zeros((B, n_full, D), dtype=sc.dtype, device=sc.device)
        else:
            dst = torch.empty((B, n, D), dtype=sc.dtype, device=sc.device)
    _lib.call("vpt_token_gather", _p(sc), _p(idx), _p(dst), B, n_full, n, D, int(scatter), _stream())
    return dst


class TokenGatherFn(torch.autograd.Function):
    """out[b, j] = x[b, idx[j]] for a [B, L, D] token buffer (TREAD: keep / route subsets of one random permutation)."""

    @staticmethod
    def forward(ctx, x3, idx):
        ctx.save_for_backward(idx)
        ctx.n_full = x3.shape[1]
        return _token_gather_call(x3, idx, x3.shape[1], False)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        return _token_gather_call(dy, idx, ctx.n_full, True), None


class TokenScatterFn(torch.autograd.Function):
    """out[b, idx[j]] = x[b, j]; rows of out not named by idx are zero.  idx must hold unique positions."""

    @staticmethod
    def forward(ctx, x3, idx, n_full):
        ctx.save_for_backward(idx)
        return _token_gather_call(x3, idx, n_full, True)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        return _token_gather_call(dy, idx, dy.shape[1], False), None, None


class TokenMergeFn(torch.autograd.Function):
    """out[b, idx_a[j]] = a[b, j], out[b, idx_b[j]] = b[b, j]: two disjoint index sets that together cover every row."""

    @staticmethod
    def forward(ctx, a3, b3, idx_a, idx_b):
        L = idx_a.numel() + idx_b.numel()
        out = torch.empty((a3.shape[0], L, a3.shape[2]), dtype=a3.dtype, device=a3.device)
        _token_gather_call(a3, idx_a, L, True, out)
        _token_gather_call(b3, idx_b, L, True, out)
        ctx.save_for_backward(idx_a, idx_b)
        return out

    @staticmethod
    def backward(ctx, dy):
        idx_a, idx_b = ctx.saved_tensors
        L = dy.shape[1]
        return _token_gather_call(dy, idx_a, L, False), _token_gather_call(dy, idx_b, L, False), None, None


def token_merge(a3, b3, idx_a, idx_b):
    return TokenMergeFn.apply(a3, b3, _check_tokens(a3, idx_a), _check_tokens(b3, idx_b))


def _check_tokens(x3, idx):
    _need_cuda(x3, idx)
    if x3.dim() != 3 or x3.element_size() != 2 or x3.shape[-1] % 8 != 0:
        raise ValueError("token gather / scatter: 2-byte [B, L, D] tensors with D a multiple of 8")
    return idx.to(device=x3.device, dtype=torch.int64).contiguous()


def token_gather(x3, idx):
    return TokenGatherFn.apply(x3, _check_tokens(x3, idx))


def token_scatter(x3, idx, n_full: int):
    return TokenScatterFn.apply(x3, _check_tokens(x3, idx), int(n_full))
