#!/usr/bin/env python
"""Headline benchmark: JiT NF4-QLoRA training images/sec on N B200s (BASELINE.json `metric`).

  python bench.py --gpus 1 --steps 20 --warmup 5                       # our arm, one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N ...                                             # data parallel, one rank per GPU
  python bench.py --impl reference ...                                  # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  A "step" is one optimisation step of the named workload: noise + forward + loss +
backward + LoRA-gradient all-reduce (N > 1) + clip + AdamW, on synthetic images of the named shape with random-init
weights.  `value` times K CUDA-graph replays with the batch already in HBM; `e2e` times the same K steps through the
public API with the batch in pinned host memory (H2D copy of the batch and D2H read of the loss inside the timed
region).  `roofline` is the fused NF4-LoRA GEMM (forward + backward-dX launches of the block linears) timed with CUDA
events inside real steps.  `cpu_baseline` is the oracle's CPU restatement of the same step on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "JiT NF4-QLoRA train images/sec"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None, help="JiT-B/16 | JiT-L/16 | JiT-H/16 (default: B at 1 GPU, L at >1)")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--rank", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=2)
    return ap.parse_args()


def workload_name(model: str, batch: int, res: int, rank: int) -> str:
    return f"{model} {res}px NF4 QLoRA (rank {rank}) bf16 training, batch {batch} per GPU"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> None:
    """The reference's CPU implementation of the path on the host cores: oracle port (the reference is Python and its
    NF4 arithmetic lives in bitsandbytes, neither of which travels to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cpu_step
    model = args.model or ("JiT-B/16" if args.gpus == 1 else "JiT-L/16")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cb = min(args.cpu_batch, args.batch)
    steps, warm = min(args.steps, 3), min(args.warmup, 1)
    r = cpu_step.time_train_steps(model=model, batch=cb, height=args.res, width=args.res, steps=steps, warmup=warm,
                                  rank=args.rank)
    v = r["images_per_s"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(model, args.batch, args.res, args.rank)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": f"{cb} images/step x {steps} steps of the same step (fp32, NF4 dequantised per call) on the host CPU"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (vision_pt_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    from vision_pt_b200 import ops
    from vision_pt_b200 import train as T

    model_name = args.model or ("JiT-B/16" if world == 1 else "JiT-L/16")
    net = T.build_jit_qlora(model_name, rank=args.rank, alpha=float(args.rank), device=dev, seed=42)
    step = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, process_group=group, use_graph=not args.no_graph,
                               seed=42 + rank)
    host = T.synthetic_batch(args.batch, args.res, args.res, seed=1000 + rank)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def load_batch():
        step.image.copy_(host[0], non_blocking=True)
        step.class_ids.copy_(host[1], non_blocking=True)
        step.attention_mask.copy_(host[2], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    load_batch()
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step.run()
    torch.cuda.synchronize()
    launches_per_step = step.kernel_launches

    # ---- device-resident throughput
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step.run()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    value = world * args.batch * args.steps / (ms_total * 1e-3)
    final_loss = float(step.loss.item())

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H loss, every step
    barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        load_batch()
        loss = step.run()
        _ = float(loss.item())          # D2H read of the step's result (synchronises, as a logging trainer would)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * args.batch * args.steps / (float(ms2.item()) * 1e-3)

    # ---- roofline of the dominant kernel: every fused-linear launch of 2 eager steps bracketed by CUDA events
    roof = None
    if rank == 0:
        ops.GEMM_TIMER = []
        was = step.use_graph
        step.use_graph = False
        for _ in range(2):
            step.run()
        torch.cuda.synchronize()
        step.use_graph = was
        recs, ops.GEMM_TIMER = ops.GEMM_TIMER, None
        dom = [r for r in recs if r["nf4"] and r["lora"]]
        t_ms = sum(r["e0"].elapsed_time(r["e1"]) for r in dom)
        fl = sum(r["flops"] for r in dom)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
        if not peak:
            peak, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        achieved = fl / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
        all_ms = sum(r["e0"].elapsed_time(r["e1"]) for r in recs)
        roof = {"bound": "tensor", "kernel": "gemm_nf4lora_kernel (NF4 + LoRA, fwd and bwd-dX launches of the block linears)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "launches": len(dom) // 2, "avg_launch_us": 1e3 * t_ms / max(len(dom), 1),
                "gemm_ms_per_step": all_ms / 2, "flops_per_step": fl / 2}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N == 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_step
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        r = cpu_step.time_train_steps(model=model_name if model_name in cpu_step.JIT_CONFIGS else "JiT-B/16",
                                      batch=args.cpu_batch, height=args.res, width=args.res, steps=args.cpu_steps, warmup=1,
                                      rank=args.rank)
        cpu = {"value": r["images_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{args.cpu_batch} images/step x {args.cpu_steps} steps of the same step (fp32, NF4 dequantised per call)"}

    if rank == 0:
        fl = T.step_flops(net.config, args.batch, args.res, args.res, rank=args.rank)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(model_name, args.batch, args.res, args.rank),
                       "global_batch": world * args.batch, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph, "gradient_checkpointing": False,
                       "l2": "no flush: one step streams far more than the 126 MB L2 (saved activations of every block)",
                       "optimizer": "AdamW over the flat LoRA buffer, clip_grad_norm 1.0", "loss": step.hp.loss_target},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roof,
            "cpu_baseline": cpu,
            "step_tflops_algorithmic": fl["total"] / 1e12,
            "achieved_tflops_step": world * fl["total"] / (ms_step * 1e-3) / 1e12,
            "final_loss": final_loss,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
