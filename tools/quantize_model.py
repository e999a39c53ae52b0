r"""NF4-prequantise a safetensors checkpoint (replacement for the reference's stale tools/quantize_model.py:7, which
is tied to one model class): works on the state dict, so any checkpoint whose Linear weights are `<path>.weight`
tensors can be prequantised and later loaded with `replace_by_prequantized_weights` + `load_state_dict` -- by this
package (vision_pt_b200.modules.quant) or by the reference (src/modules/quant/functional.py:332-339; the key set is
bitsandbytes' `QuantState.as_dict(packed=True)`: weight, weight.absmax, weight.quant_map, weight.nested_absmax,
weight.nested_quant_map, weight.quant_state.bitsandbytes__nf4).

  python tools/quantize_model.py model.safetensors model.bnb_nf4.safetensors \
      --include 'blocks\.\d+\.(attn|mlp)\..*\.weight$' --exclude lora_ --verify

`--include` / `--exclude` are regular expressions on the tensor names (2-D floating-point tensors only are eligible).
Needs a B200: quantisation runs through the library's `vpt_nf4_quantize`."""
from __future__ import annotations

import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def select_keys(state_dict: dict[str, torch.Tensor], include: list[str], exclude: list[str]) -> list[str]:
    inc, exc = [re.compile(p) for p in include], [re.compile(p) for p in exclude]
    out = []
    for k, v in state_dict.items():
        if v.dim() != 2 or not v.dtype.is_floating_point or v.numel() % 64 != 0:
            continue
        if inc and not any(p.search(k) for p in inc):
            continue
        if any(p.search(k) for p in exc):
            continue
        out.append(k)
    return out


def quantize_file(src: str, dst: str, include: list[str], exclude: list[str], verify: bool = False) -> dict:
    from safetensors.torch import load_file, save_file

    from vision_pt_b200 import ops
    from vision_pt_b200.modules.quant import _unpack_meta, quantize_state_dict
    from vision_pt_b200.modules.state_dict import RegexMatch
    sd = load_file(src)
    keys = select_keys(sd, include, exclude)
    if not keys:
        raise ValueError("no tensor matches --include / --exclude")
    originals = {k: sd[k].clone() for k in keys} if verify else {}
    before = sum(v.numel() * v.element_size() for v in sd.values())
    # exact names (get_target_keys treats plain strings as substrings)
    quantize_state_dict(sd, "bnb_nf4", [RegexMatch(regex=re.escape(k) + "$") for k in keys])
    after = sum(v.numel() * v.element_size() for v in sd.values())
    report = {"quantized": len(keys), "bytes_before": before, "bytes_after": after, "max_abs_err": None}
    if verify:
        worst = 0.0
        for k in keys:
            meta = _unpack_meta(sd[f"{k}.quant_state.bitsandbytes__nf4"])
            st = ops.Nf4Tensors(packed=sd[k].cuda(), absmax=sd[f"{k}.absmax"].cuda(), nested_absmax=sd[f"{k}.nested_absmax"].cuda(),
                                nested_code=sd[f"{k}.nested_quant_map"].cuda(), code=sd[f"{k}.quant_map"].cuda(),
                                offset=float(meta["nested_offset"]), shape=tuple(meta["shape"]),
                                dtype=getattr(torch, meta["dtype"]))
            w = ops.nf4_dequantize(st).float().cpu()
            ref = originals[k].float()
            worst = max(worst, float((w - ref).abs().max() / ref.abs().max().clamp_min(1e-12)))
        report["max_abs_err"] = worst     # relative to the tensor's absolute maximum; NF4's coarsest step is ~0.14
    save_file({k: v.contiguous() for k, v in sd.items()}, dst, metadata={"format": "pt", "quant_type": "bnb_nf4"})
    return report


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--include", action="append", default=[], help="regex on tensor names (repeatable); default: every 2-D weight")
    ap.add_argument("--exclude", action="append", default=[], help="regex on tensor names to leave alone (repeatable)")
    ap.add_argument("--verify", action="store_true", help="dequantise again and report the worst relative error")
    args = ap.parse_args(argv)
    rep = quantize_file(args.src, args.dst, args.include, args.exclude, args.verify)
    print(f"{rep['quantized']} tensors -> NF4, {rep['bytes_before'] / 2**20:.1f} MiB -> {rep['bytes_after'] / 2**20:.1f} MiB"
          + (f", worst error {rep['max_abs_err']:.3f} of the tensor maximum" if rep["max_abs_err"] is not None else ""))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
