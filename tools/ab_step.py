"""Interleaved A/B timing of whole training steps that differ in one switch (same process, same weights, same thermal state).

Every variant is its own CUDA graph captured under its switch setting over ONE shared training state; the timed loops
alternate (A x n, B x n, ... repeated `--rounds` times) so that the power-capped clock drift hits all variants alike.
  python tools/ab_step.py [--model JiT-B/16] [--batch 64] [--res 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import ops  # noqa: E402
from vision_pt_b200 import train as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="JiT-B/16")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--rounds", type=int, default=4)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()

VARIANTS = {
    "default": {},
    "separate swiglu kernels": {"FUSE_SWIGLU": False},
    "no dequant prefetch": {"PREFETCH_DEQUANT": False},
    "separate q / k / v GEMMs": {"FUSE_QKV": False},
    "ATen glue (noise, slices, dense final layer)": {"FUSED_GLUE": False},
}
net = T.build_jit_qlora(args.model, device="cuda", seed=42)
state = T.TrainState(net)
host = T.synthetic_batch(args.batch, args.res, args.res)
steps = {}
for name, sw in VARIANTS.items():
    saved = {k: getattr(ops, k) for k in sw}
    for k, v in sw.items():
        setattr(ops, k, v)
    st = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, state=state, seed=None)
    st.image.copy_(host[0]); st.class_ids.copy_(host[1]); st.attention_mask.copy_(host[2])
    st.capture()
    for k, v in saved.items():
        setattr(ops, k, v)
    steps[name] = st
for _ in range(60):                       # settle into the power-capped steady state
    for st in steps.values():
        st.run()
torch.cuda.synchronize()
acc = {n: [] for n in steps}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(args.rounds):
    for name, st in steps.items():
        e0.record()
        for _ in range(args.steps):
            st.run()
        e1.record()
        torch.cuda.synchronize()
        acc[name].append(e0.elapsed_time(e1) / args.steps)
base = sum(acc["default"]) / len(acc["default"])
print(f"# {args.model} batch {args.batch} {args.res}px, {args.rounds} interleaved rounds x {args.steps} steps, ms/step (mean; per round)")
for name, v in acc.items():
    m = sum(v) / len(v)
    print(f"{name:40s} {m:8.3f}  ({m / base - 1:+.1%})   " + " ".join(f"{x:.3f}" for x in v))
