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
    ap.add_argument("--model", default="JiT-B/16", help="JiT-B/16 | JiT-L/16 (the headline workload is JiT-B/16 at every N so that N = 1, 2, 4, 8 measure the same per-GPU work)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary JiT-L/16 data-parallel measurement at N > 1")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--rank", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--checkpointing", action="store_true", help="gradient checkpointing per block (the reference's shipped YAML sets it)")
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "radam_schedulefree"],
                    help="update rule of the timed step (the shipped YAMLs name schedulefree.RAdamScheduleFree)")
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


# ------------------------------------------------------------------------------------------------ roofline
def gemm_roofline(step, ops, torch, replay_here: bool = True):
    ops.GEMM_TIMER = []
    was = step.use_graph
    step.use_graph = False
    step.run()
    torch.cuda.synchronize()
    step.use_graph = was
    recs, ops.GEMM_TIMER = ops.GEMM_TIMER, None
    tape = [r for r in recs if r["nf4"] and r["lora"] and r["scratch"]]
    if not tape or not replay_here:
        return None
    dev = torch.device("cuda", torch.cuda.current_device())
    pool: dict = {}

    def buf(shape, stride, salt):
        """A few distinct buffers per shape, rotated, so that consecutive calls do not find their input in L2."""
        key = (tuple(shape), stride)
        lst = pool.setdefault(key, [])
        if len(lst) < 3:
            t = torch.randn((shape[0], stride), device=dev).mul_(0.5).to(torch.bfloat16)[:, :shape[1]]
            lst.append(t)
        return lst[salt % len(lst)]

    def replay(reuse: bool):
        outs = []
        for i, r in enumerate(tape):
            shape, stride, w, bias, down, up, scale, has_res, want_side, backward = r["call"]
            n_out = w.shape[1] if backward else w.shape[0]
            res = buf((shape[0], n_out), (n_out + 7) // 8 * 8, i + 1) if has_res else None
            outs.append(ops.linear_raw(buf(shape, stride, i), w, bias, down, up, scale, res, want_side=want_side,
                                       backward=backward, reuse_scratch=reuse))
        return outs

    def timed(reuse: bool) -> float:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            replay(reuse)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = replay(reuse)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del keep, g
        return best

    ms_gemm = timed(True)
    ms_call = timed(False)
    fl = sum(r["flops"] for r in tape)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long back-to-back sequence)"
    if not peak:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_pair_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    achieved = fl / (ms_gemm * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "gemm_pair_kernel (NF4 + LoRA linears of the blocks, forward and backward-dX)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "launches": len(tape), "avg_launch_us": 1e3 * ms_gemm / len(tape),
            "flops_per_launch": fl / len(tape),
            "with_dequant": {"avg_call_us": 1e3 * ms_call / len(tape), "achieved": fl / (ms_call * 1e-3) / 1e12,
                             "note": "same calls including the per-call NF4 dequantisation kernel (HBM-bound, 0 FLOP counted)"},
            "how": "CUDA-graph replay of the step's taped calls, inputs rotated over 3 buffers per shape, best of 3"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args) -> None:
    """The reference's CPU implementation of the path on the host cores: oracle port (the reference is Python and its
    NF4 arithmetic lives in bitsandbytes, neither of which travels to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cpu_step
    model = args.model
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
    if os.environ.get("VPT_DIST_BACKEND", "nccl") != "nccl":
        local = 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (vision_pt_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        backend = os.environ.get("VPT_DIST_BACKEND", "nccl")     # "gloo" only to exercise the N > 1 path on a 1-GPU box
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group(backend)
        group = dist.group.WORLD

    from vision_pt_b200 import ops
    from vision_pt_b200 import train as T

    model_name = args.model
    net = T.build_jit_qlora(model_name, rank=args.rank, alpha=float(args.rank), device=dev, seed=42)
    net.set_gradient_checkpointing(args.checkpointing)
    hp = T.TrainHParams(optimizer=args.optimizer)
    step = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, process_group=group, use_graph=not args.no_graph,
                               seed=42 + rank, hp=hp)
    host = T.synthetic_batch(args.batch, args.res, args.res, seed=1000 + rank)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def load_batch():
        step.image.copy_(host[0], non_blocking=True)
        step.class_ids.copy_(host[1], non_blocking=True)
        step.attention_mask.copy_(host[2], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local]) if dist.get_backend() == "nccl" else dist.barrier()

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
    loss_target = step.hp.loss_target

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

    # ---- roofline of the dominant kernel (the CTA-pair NF4-LoRA GEMM): the fused-linear calls of one real step are taped,
    # then replayed back to back from a CUDA graph on buffers of the same shapes (inputs rotate over > L2 worth of memory)
    # and timed with CUDA events on the launching stream -- once GEMM only (`reuse_scratch`: the dequantised weight is
    # already in the workspace) and once as issued in the step (dequantisation + GEMM).
    # Every rank runs the taping step (it contains the gradient all-reduce); only rank 0 replays and reports.
    roof = gemm_roofline(step, ops, torch, replay_here=rank == 0)
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

    # ---- secondary workload at N > 1: BASELINE.json configs[2], JiT-L/16 data parallel (same step, same timing rules)
    extra = None
    if world > 1 and not args.no_extra and model_name != "JiT-L/16":
        del step, net
        torch.cuda.empty_cache()
        net = T.build_jit_qlora("JiT-L/16", rank=args.rank, alpha=float(args.rank), device=dev, seed=42)
        step = T.JiTQLoRATrainStep(net, args.batch, args.res, args.res, process_group=group, use_graph=not args.no_graph,
                                   seed=42 + rank, hp=hp)
        load_batch_l = lambda: (step.image.copy_(host[0], non_blocking=True), step.class_ids.copy_(host[1], non_blocking=True),
                                step.attention_mask.copy_(host[2], non_blocking=True))
        load_batch_l()
        for _ in range(3):
            step.run()
        ksteps = max(3, args.steps // 2)
        barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(ksteps):
            step.run()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        msl = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(msl, op=dist.ReduceOp.MAX)
        fl_l = T.step_flops(net.config, args.batch, args.res, args.res, rank=args.rank)
        extra = {"workload": workload_name("JiT-L/16", args.batch, args.res, args.rank), "value": world * args.batch * ksteps / (float(msl.item()) * 1e-3),
                 "unit": UNIT, "ms_per_step": float(msl.item()) / ksteps, "steps": ksteps,
                 "achieved_tflops_step": world * fl_l["total"] / (float(msl.item()) / ksteps * 1e-3) / 1e12}
        net_cfg_for_flops = T.MODEL_CONFIGS[model_name]()
    else:
        net_cfg_for_flops = net.config

    if rank == 0:
        fl = T.step_flops(net_cfg_for_flops, args.batch, args.res, args.res, rank=args.rank)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(model_name, args.batch, args.res, args.rank),
                       "global_batch": world * args.batch, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph, "gradient_checkpointing": bool(args.checkpointing),
                       "l2": "no flush: one step streams far more than the 126 MB L2 (saved activations of every block)",
                       "optimizer": ("AdamW" if args.optimizer == "adamw" else "schedulefree.RAdamScheduleFree")
                                    + " over the flat LoRA buffer, clip_grad_norm 1.0", "loss": loss_target},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roof,
            "cpu_baseline": cpu,
            "extra_workload": extra,
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
