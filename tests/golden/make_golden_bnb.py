"""The ONE command that pins oracle/nf4.py to bitsandbytes itself, for the day a box with the wheel exists:

    pip install bitsandbytes==0.48.2          # the reference's pin, /root/reference/uv.lock:308-309
    python tests/golden/make_golden_bnb.py    # needs a CUDA device; writes tests/golden/bnb_nf4_vectors.pt (~3 MB)

It quantises seeded bf16 weights of the hot-path shapes with `bitsandbytes.functional.quantize_4bit(blocksize=64,
compress_statistics=True, quant_type="nf4")`, dequantises them with `dequantize_4bit`, and stores the packed codes, the
statistics and a strided sample of the dequantised values.  tests/test_oracle_nf4_bnb.py (skipped while the file is absent)
then requires oracle/nf4.py to reproduce every byte.  Neither this image nor the GPU box has the wheel (no network), which is
why the NF4 arithmetic is "parity unpinned" today.
"""
import os
import sys

import torch

try:
    import bitsandbytes.functional as BF
except ImportError:
    sys.exit("bitsandbytes is not installed here: nothing written (see the docstring)")

SHAPES = [(768, 768), (2048, 768), (768, 2048), (1024, 2730), (2730, 1024), (1280, 3413), (3413, 1280), (640, 2048)]
out = {}
for i, (n, k) in enumerate(SHAPES):
    g = torch.Generator().manual_seed(100 + i)
    w = (torch.randn(n, k, generator=g) * 0.02).to(torch.bfloat16).cuda()
    packed, st = BF.quantize_4bit(w, blocksize=64, compress_statistics=True, quant_type="nf4", quant_storage=torch.uint8)
    deq = BF.dequantize_4bit(packed, st)
    d = st.as_dict(packed=True)
    out[f"{n}x{k}"] = {"weight": w.cpu(), "packed": packed.cpu(), "stats": {a: b.cpu() if torch.is_tensor(b) else b for a, b in d.items()},
                       "dequant_sample": deq.reshape(-1)[::97].cpu().clone(), "dequant_sum": float(deq.float().sum()),
                       "dequant_first_rows": deq[:4].cpu().clone()}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bnb_nf4_vectors.pt")
torch.save(out, path)
print("wrote", path)
