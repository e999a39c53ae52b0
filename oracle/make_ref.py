"""Recipe for oracle/_ref: the reference's OWN Python modules of the hot path, staged where they can travel to the GPU box.
TEST INFRASTRUCTURE ONLY (see oracle/nf4.py for the import rule).

    python oracle/make_ref.py            # copies /root/reference/src/**/*.py  ->  oracle/_ref/src/

The reference is pure Python, so "building" it is staging its source files next to the oracle: `oracle/_ref/` is listed in
.gitignore (nothing of the reference enters the history) but not in .gpurunignore, so the files ride along with the built
.so to the GPU box, where /root/reference does not exist.  __graft_entry__.build() runs this whenever /root/reference is
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


def make_ref(force: bool = False) -> str | None:
    """Returns the staged root, or None when the reference is not available (GPU box: the staged copy is already there)."""
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        return OUT if os.path.isdir(os.path.join(OUT, "src")) else None
    stamp = os.path.join(OUT, ".stamp")
    newest = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= newest:
        return OUT
    dst = os.path.join(OUT, "src")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    n = 0
    for d, _, fs in os.walk(src):
        for f in fs:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(d, f), src)
            to = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(to), exist_ok=True)
            shutil.copyfile(os.path.join(d, f), to)
            n += 1
    with open(stamp, "w") as fh:
        fh.write(f"{n} files staged from {src}\n")
    return OUT


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
