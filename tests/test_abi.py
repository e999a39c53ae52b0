"""The C-ABI library builds, loads and exports every symbol include/vptb200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vptb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_built):
    lib = ctypes.CDLL(lib_built)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vptb200.h but not exported"


def test_python_binding_covers_header(lib_built):
    from vision_pt_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"vpt_last_error", "vpt_abi_version", "vpt_linear_scratch_bytes", "vpt_linear_scratch_bytes_dir"}
    assert bound == set(_declared())
    lib = _lib.load()
    assert lib.vpt_abi_version() == 7
    assert lib.vpt_linear_scratch_bytes(768, 2048) >= 2 * 768 * 2048


def test_argument_errors_are_reported_without_a_gpu(lib_built):
    """Validation happens before any CUDA call: bad arguments give a non-zero code and a message."""
    from vision_pt_b200 import _lib
    lib = _lib.load()
    rc = lib.vpt_rmsnorm_fwd(None, None, None, None, 0, 0, 0, 0, 0.0, None)
    assert rc != 0 and b"vpt_rmsnorm_fwd" in lib.vpt_last_error()
    args = _lib.LinearArgsC()
    rc = lib.vpt_nf4lora_linear_fwd(ctypes.byref(args), None)
    assert rc != 0


def test_no_oracle_import_in_product():
    """The product package never touches oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "vision_pt_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(d, f)).read(), os.path.join(d, f)
