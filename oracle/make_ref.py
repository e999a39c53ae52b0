"""Recipe for oracle/_ref: the reference's OWN Python modules of the hot path, staged where they can travel to the GPU box.
TEST INFRASTRUCTURE ONLY (see oracle/nf4.py for the import rule).

    python oracle/make_ref.py            # packs /root/reference/src/**/*.py  ->  oracle/_ref/reference_src.tar.gz

The reference is pure Python, so "building" it is packing its modules into ONE archive next to the oracle (the counterpart
of the .so a compiled reference would leave there): `oracle/_ref/` is listed in .gitignore (nothing of the reference enters
the history) but not in .gpurunignore, so the archive rides along with the built .so to the GPU box, where /root/reference
does not exist; oracle/refimport.py unpacks it into a temporary directory at import time.  __graft_entry__.build() runs this whenever /root/reference is
present.  oracle/refimport.py imports the staged modules (package __init__ files that pull accelerate / bitsandbytes are
bypassed the way SURVEY.md section 8c describes); oracle/ref_runner.py times the reference's Denoiser / LoRALinear /
scaled_dot_product_attention with them.  Nothing under vision_pt_b200/ reads this directory.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VPT_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


ARCHIVE = os.path.join(OUT, "reference_src.tar.gz")


def make_ref(force: bool = False) -> str | None:
    """Returns the archive path, or None when neither the reference nor a packed copy is available."""
    import tarfile
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    newest = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
    if not force and os.path.exists(ARCHIVE) and os.path.getmtime(ARCHIVE) >= newest:
        return ARCHIVE
    os.makedirs(OUT, exist_ok=True)
    loose = os.path.join(OUT, "src")                   # an earlier layout kept loose files: gone
    if os.path.isdir(loose):
        shutil.rmtree(loose)
    n = 0
    with tarfile.open(ARCHIVE + ".tmp", "w:gz") as tar:
        for d, _, fs in sorted(os.walk(src)):
            for f in sorted(fs):
                if f.endswith(".py"):
                    full = os.path.join(d, f)
                    tar.add(full, arcname=os.path.join("src", os.path.relpath(full, src)))
                    n += 1
    os.replace(ARCHIVE + ".tmp", ARCHIVE)
    with open(os.path.join(OUT, ".stamp"), "w") as fh:
        fh.write(f"{n} files packed from {src}\n")
    return ARCHIVE


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
