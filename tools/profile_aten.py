"""Which kernels of one eager training step are NOT ours?  torch.profiler over one step: every CUDA kernel that is not a
libvptb200 kernel, grouped by name, with the input shapes of the ATen op that launched it (where the copies / fills of the
step come from).  Usage: python tools/profile_aten.py [--model JiT-B/16] [--batch 64]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import train as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="JiT-B/16")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--res", type=int, default=256)
args = ap.parse_args()

net = T.build_jit_qlora(args.model, device="cuda", seed=42)
step = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, use_graph=False)
host = T.synthetic_batch(args.batch, args.res, args.res)
step.image.copy_(host[0]); step.class_ids.copy_(host[1]); step.attention_mask.copy_(host[2])
for _ in range(2):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step.run()
    torch.cuda.synchronize()

ours_us = other_us = 0.0
rows = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    if "vpt::" in ev.name:
        ours_us += dur
        continue
    other_us += dur
    rows[ev.name[:90]][0] += 1
    rows[ev.name[:90]][1] += dur
print(f"ours {ours_us / 1e3:.3f} ms, other {other_us / 1e3:.3f} ms")
for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{us:9.1f} us  n={n:4d}  {name}")
print("\nATen ops with CUDA time (shapes):")
agg = prof.key_averages(group_by_input_shape=True)
lines = []
for a in agg:
    t = getattr(a, "self_device_time_total", None)
    if t is None:
        t = a.self_cuda_time_total
    if t > 0 and a.key.startswith("aten::"):
        lines.append((t, a.count, a.key, str(a.input_shapes)[:110]))
for t, n, k, shp in sorted(lines, reverse=True)[:30]:
    print(f"{t:9.1f} us  n={n:4d}  {k:28s} {shp}")
