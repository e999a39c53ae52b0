"""Aspect-ratio-bucketed training throughput (BASELINE.json configs[3]: JiT-H/16 512 px, NF4 QLoRA, (H, W) drawn per step
and per rank from the buckets of src/dataset/aspect_ratio_bucket.py:20-60 with base 512, step 64, min 256), through
`JiTQLoRATrainer.train_step` -- the call a user makes: pinned host batch -> H2D -> the bucket's CUDA graph -> loss on the
device.  One graph per bucket over one shared LoRA / optimiser state.

  python tools/bench_arb.py [--model JiT-H/16 --batch 16 --steps 36]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_arb.py   (data parallel)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_pt_b200 import train as T  # noqa: E402


def buckets(base: int = 512, step: int = 64, lo: int = 256) -> list[tuple[int, int]]:
    """The reference's rule (generate_buckets, src/dataset/aspect_ratio_bucket.py:20-60): one side walks down from `base` in
    `step`s while it is >= `lo`; the other is base^2 / side rounded to the nearest multiple of `step` (the walk stops when
    that falls below `lo`); every pair and its transpose is a bucket.  base 512 / step 64 / min 256 -> 9 buckets:
    512x512, 448x576, 384x704, 320x832, 256x1024 and transposes (1008..1056 patches of 16 px)."""
    out = []
    side = base
    while side >= lo:
        other = round(base * base / side / step) * step
        if other < lo:
            break
        for hw in ((side, other), (other, side)):
            if hw not in out:
                out.append(hw)
        side -= step
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="JiT-H/16")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=36)
    ap.add_argument("--optimizer", default="adamw")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    net = T.build_jit_qlora(args.model, device=dev, seed=42)
    tr = T.JiTQLoRATrainer(net, hp=T.TrainHParams(optimizer=args.optimizer), process_group=group, seed=42 + rank)
    bks = buckets()
    host = {hw: T.synthetic_batch(args.batch, hw[0], hw[1], seed=7 + i) for i, hw in enumerate(bks)}
    # capture every bucket's graph up front, all ranks in lock-step: the LoRA-gradient all-reduce is then part of each graph
    tr.precapture([(args.batch, hw[0], hw[1]) for hw in bks])
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(1000 + rank)   # every rank draws its own bucket per step
    order = [bks[int(torch.randint(len(bks), (1,), generator=g))] for _ in range(args.steps)]
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for hw in order:
        loss = tr.train_step(*host[hw])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per_bucket = {}
    for hw in bks:                                   # per-bucket step time (device-resident batch, graph replay only)
        st = tr.bucket(args.batch, hw[0], hw[1])
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            st.run()
        e1.record()
        torch.cuda.synchronize()
        per_bucket[f"{hw[0]}x{hw[1]}"] = round(e0.elapsed_time(e1) / 3, 2)
    if rank == 0:
        tokens = {f"{h}x{w}": (h // 16) * (w // 16) + 10 + 64 for h, w in bks}
        print(json.dumps({"metric": "JiT NF4-QLoRA train images/sec (aspect-ratio buckets)", "unit": "images/s",
                          "value": world * args.batch * args.steps / (float(ms.item()) * 1e-3), "n_gpus": world,
                          "ms_per_step": float(ms.item()) / args.steps, "steps": args.steps, "final_loss": float(loss),
                          "config": {"workload": f"{args.model} NF4 QLoRA rank 16, batch {args.batch} per GPU, buckets of base 512 / step 64 / min 256",
                                     "buckets": len(bks), "tokens_per_sample": tokens, "optimizer": args.optimizer},
                          "ms_per_step_by_bucket": per_bucket, "graphs": len(tr.buckets)}))
    if world > 1:
        # graphs that hold the captured all-reduce go first; tearing the communicator down under them blocks (trainer.close)
        tr.close()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
