"""Why is e2e below value?  Alternates the two loops and switches pieces of the e2e loop off."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vision_pt_b200 import train as T  # noqa: E402

net = T.build_jit_qlora("JiT-B/16", rank=16, alpha=16.0, device="cuda", seed=42)
tr = T.JiTQLoRATrainer(net, seed=42)
step = tr.bucket(64, 256, 256)
hosts = [T.synthetic_batch(64, 256, 256, seed=1000 + 7919 * i) for i in range(2)]
tr.train_step(*hosts[0])
for _ in range(5):
    step.run()
torch.cuda.synchronize()


def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def v_plain(i):
    step.run()


def v_e2e(i):
    tr.train_step(*hosts[i % 2], prefetch=hosts[(i + 1) % 2])
    if i > 0:
        tr.read_loss(1)


def v_noprefetch(i):            # direct H2D on the main stream, no loss read
    tr.train_step(*hosts[i % 2])


def v_prefetch_noread(i):
    tr.train_step(*hosts[i % 2], prefetch=hosts[(i + 1) % 2])


def v_d2d_only(i):              # D2D copy + replay, nothing from the host
    step.image.copy_(step.image, non_blocking=True) if False else None
    step.run()
    tr._loss_ring[0:1].copy_(step.loss.reshape(1), non_blocking=True)


tr.prefetch(*hosts[0])
for name, fn in (("value", v_plain), ("e2e", v_e2e), ("value", v_plain), ("e2e", v_e2e), ("prefetch, no loss read", v_prefetch_noread),
                 ("no prefetch (H2D on main stream)", v_noprefetch), ("replay + D2H loss copy", v_d2d_only), ("value", v_plain),
                 ("value x100", lambda i: step.run())):
    n = 100 if "x100" in name else 20
    if "e2e" in name or "prefetch" in name:
        tr._staging.clear()
        tr.prefetch(*hosts[0])
    print(f"{name:36s} {timed(fn, n):8.3f} ms/step")
