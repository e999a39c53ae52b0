"""ctypes binding of libvptb200.so (include/vptb200.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPT_LIB") or os.path.join(_HERE, "libvptb200.so")   # VPT_LIB: A/B builds of the same library

VPT_BF16, VPT_F16, VPT_F32 = 0, 1, 2


class Nf4WeightC(C.Structure):
    _fields_ = [
        ("packed", C.c_void_p), ("qabsmax", C.c_void_p), ("nested_absmax", C.c_void_p), ("nested_code", C.c_void_p),
        ("code", C.c_void_p), ("offset", C.c_float), ("N", C.c_int32), ("K", C.c_int32),
        ("packed_rows", C.c_void_p), ("absmax_f32", C.c_void_p), ("K_pad", C.c_int32),
    ]


class LinearArgsC(C.Structure):
    _fields_ = [
        ("w", Nf4WeightC), ("w_bf16", C.c_void_p), ("bias", C.c_void_p), ("lora_down", C.c_void_p),
        ("ld_lora_down", C.c_int64), ("lora_up", C.c_void_p), ("scale", C.c_float), ("inp", C.c_void_p), ("ld_in", C.c_int64), ("out", C.c_void_p),
        ("ld_out", C.c_int64), ("residual", C.c_void_p), ("ld_res", C.c_int64), ("side", C.c_void_p), ("M", C.c_int32),
        ("tile_n", C.c_int32), ("w_scratch", C.c_void_p), ("ld_scratch", C.c_int64), ("scratch_bytes", C.c_int64), ("ld_side", C.c_int64), ("reuse_scratch", C.c_int32),
        ("epilogue", C.c_int32), ("in2", C.c_void_p), ("ld_in2", C.c_int64), ("out2", C.c_void_p), ("ld_out2", C.c_int64),
        ("n_sections", C.c_int32),
    ]


class LoraGradItemC(C.Structure):
    _fields_ = [("src", C.c_void_p), ("ld_src", C.c_int64), ("M", C.c_int32), ("P", C.c_int32), ("nsmall", C.c_int32),
                ("small_t", C.c_void_p * 3), ("ld_small", C.c_int64), ("out", C.c_void_p * 3), ("transposed", C.c_int32),
                ("ld_out", C.c_int64)]


class DequantItemC(C.Structure):
    _fields_ = [("w", Nf4WeightC), ("w_scratch", C.c_void_p), ("scratch_bytes", C.c_int64), ("lora_down", C.c_void_p),
                ("ld_lora_down", C.c_int64), ("lora_up", C.c_void_p)]


class AttnTensorC(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sl", C.c_int64), ("sh", C.c_int64)]


_P, _I32, _I64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_AT = C.POINTER(AttnTensorC)

# name -> argtypes (the stream is always last); every function returns int
SIGNATURES: dict[str, list] = {
    "vpt_nf4_dequant": [C.POINTER(Nf4WeightC), _I64, C.c_int, _P, _P],
    "vpt_nf4_repack": [C.POINTER(Nf4WeightC), _P, _P, _I32, _P],
    "vpt_nf4_quantize": [_P, C.c_int, _I64, _P, _P, _P, _P, _P, _P, _P],
    "vpt_nf4lora_linear_fwd": [C.POINTER(LinearArgsC), _P],
    "vpt_nf4lora_linear_bwd_dx": [C.POINTER(LinearArgsC), _P],
    "vpt_lora_grad_batch": [C.POINTER(LoraGradItemC), _I32, _P],
    "vpt_nf4_dequant_batch": [C.POINTER(DequantItemC), _I32, _I32, _P],
    "vpt_attn_fwd": [_AT, _AT, _AT, _AT, _I32, _I32, _I32, _I32, _I32, _P, _F, _P, _P],
    "vpt_attn_bwd": [_AT, _AT, _AT, _AT, _AT, _AT, _AT, _AT, _I32, _I32, _I32, _I32, _I32, _P, _F, _P, _P, _P],
    "vpt_rmsnorm_fwd": [_P, _P, _P, _P, _I64, _I32, _I64, _I64, _F, _P],
    "vpt_rmsnorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I64, _F, _P],
    "vpt_qknorm_rope_fwd": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I64, _I64, _F, _P],
    "vpt_qknorm_rope_bwd": [_P, _I32, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I64, _I64, _I64, _F, _P],
    "vpt_swiglu_fwd": [_P, _P, _P, _I64, _I32, _I64, _I64, _I64, _P],
    "vpt_swiglu_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _I64, _I64, _I64, _I64, _I64, _P],
    "vpt_ln_modulate_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P],
    "vpt_ln_modulate_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P],
    "vpt_gate_residual_fwd": [_P, _P, _P, _P, _I64, _I32, _I32, _P],
    "vpt_gate_residual_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P],
    "vpt_patchify": [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "vpt_unpatchify": [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P],
    "vpt_copy_rows": [_P, _I64, _P, _I64, _I64, _I64, _P],
    "vpt_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I32, _F, _P],
    "vpt_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    "vpt_gated_act_fwd": [_P, _P, _P, _I64, _I32, _I64, _I64, _I64, _I32, _P],
    "vpt_gated_act_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _I64, _I64, _I64, _I64, _I64, _I32, _P],
    "vpt_act_fwd": [_P, _P, _I64, _I32, _P],
    "vpt_act_bwd": [_P, _P, _P, _I64, _I32, _P],
    "vpt_rope_half": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _I64, _I64, _I32, _P],
    "vpt_pope_fwd": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I64, _I64, _P],
    "vpt_pope_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I64, _I64, _I64, _P],
    "vpt_token_gather": [_P, _P, _P, _I32, _I64, _I64, _I32, _I32, _P],
    "vpt_grad_sumsq": [_P, _I64, _F, _P, _P],
    "vpt_adamw_step": [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _P, _F, _P, _I32, _P],
    "vpt_radam_schedulefree_step": [_P, _P, _P, _P, _I64, _D, _D, _D, _F, _F, _D, _D, _I32, _F, _P, _F, _P, _P, _I32, _P],
    "vpt_radam_schedulefree_swap": [_P, _P, _I64, _F, _I32, _P],
    "vpt_flow_loss": [_P, _P, _P, C.c_int, _P, _I64, _I64, _I32, _F, _P, _P, _P],
    "vpt_noise_mix": [_P, _P, C.c_int, _P, _I64, _I64, _F, _I32, _P, _P, _P],
    "vpt_scale_by_scalar": [_P, _P, _P, _I64, _P],
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA library; raises if it has not been built (python vision_pt_b200/csrc/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python vision_pt_b200/csrc/build.py` "
            "(vision_pt_b200 has no CPU or eager fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.vpt_last_error.restype = C.c_char_p
    lib.vpt_last_error.argtypes = []
    lib.vpt_abi_version.restype = C.c_int
    lib.vpt_linear_scratch_bytes.restype = C.c_int64
    lib.vpt_linear_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.vpt_linear_scratch_bytes_dir.restype = C.c_int64
    lib.vpt_linear_scratch_bytes_dir.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


# kernels launched per C-ABI call (vpt_attn_bwd = delta pre-pass + main kernel); used for the bench's gpu_launches
_KERNELS_PER_CALL = {"vpt_attn_bwd": 2, "vpt_nf4_quantize": 3, "vpt_radam_schedulefree_step": 2}
CALLS: dict[str, int] = {}
_launches = 0


def launch_count() -> int:
    """Number of libvptb200 kernels launched by this process so far."""
    return _launches


def add_launches(n: int) -> None:
    """Extra kernels a call launched beyond the first (e.g. the per-call NF4 dequantisation in front of a GEMM)."""
    global _launches
    _launches += n


def call(name: str, *args) -> None:
    global _launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.vpt_last_error().decode()}")
    _launches += _KERNELS_PER_CALL.get(name, 1)
    CALLS[name] = CALLS.get(name, 0) + 1
