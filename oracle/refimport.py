"""Imports the reference's own modules of the hot path from the archive oracle/_ref/reference_src.tar.gz (packed by
oracle/make_ref.py, unpacked into the temp directory) or, in the build container, straight from /root/reference.  TEST INFRASTRUCTURE ONLY (see oracle/nf4.py for the import rule).

The package __init__ files of src.models.{jit,cogview4,sdxl} import their pipelines (accelerate, bitsandbytes, ... --
absent from this image), so those packages are pre-registered as bare namespace modules and only the files that compute
are executed.  `src.models.jit.pipeline` (needed by the extension modules for a type name only) is
stubbed.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_loaded_root: str | None = None


def _unpacked_archive() -> str | None:
    """oracle/_ref/reference_src.tar.gz (oracle/make_ref.py) unpacked once per archive version into the temp directory."""
    import hashlib
    import tarfile
    import tempfile
    arc = os.path.join(HERE, "_ref", "reference_src.tar.gz")
    if not os.path.exists(arc):
        return None
    tag = hashlib.sha1(f"{os.path.getsize(arc)}-{int(os.path.getmtime(arc))}-{os.getuid()}".encode()).hexdigest()[:12]
    root = os.path.join(tempfile.gettempdir(), f"vpt_reference_{tag}")
    if not os.path.isdir(os.path.join(root, "src", "models", "jit")):
        tmp = f"{root}.{os.getpid()}"
        with tarfile.open(arc, "r:gz") as tar:
            tar.extractall(tmp, filter="data")
        try:
            os.rename(tmp, root)
        except OSError:                       # another process unpacked it meanwhile
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    return root


def reference_root() -> str | None:
    root = _unpacked_archive()
    if root is not None:
        return root
    cand = os.environ.get("VPT_REFERENCE_ROOT", "/root/reference")
    return cand if os.path.isdir(os.path.join(cand, "src", "models", "jit")) else None


def available() -> bool:
    return reference_root() is not None


def load(root: str | None = None) -> str:
    """Makes `import src....` resolve to the reference's files.  Returns the root used."""
    global _loaded_root
    if _loaded_root is not None:
        return _loaded_root
    root = root or reference_root()
    if root is None:
        raise ImportError("the reference's modules are not staged: run `python oracle/make_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("src.models", "src.models.jit", "src.models.jit.extension", "src.models.cogview4", "src.models.sdxl"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(root, *name.split("."))]
            sys.modules[name] = m
    if "src.models.jit.pipeline" not in sys.modules:
        stub = types.ModuleType("src.models.jit.pipeline")

        class JiTModel:  # only subclassed by the extension modules' model wrappers, which the oracle never instantiates
            pass

        stub.JiTModel = JiTModel
        sys.modules["src.models.jit.pipeline"] = stub
    _loaded_root = root
    return root
