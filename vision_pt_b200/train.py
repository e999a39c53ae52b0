"""JiT NF4-QLoRA training step on the sm_100a kernels, one process per GPU.

Mirrors the reference's step for this path -- train/jit/class_to_image.py:166-243 (`train_step`: class-label context,
timestep sampling, noising, denoiser call, `treat_loss`), src/trainer/common.py:326-388 (backward, gradient sync,
optimizer step, zero_grad) and src/models/for_training.py:98-109 (gradient-norm clipping) -- with the B200-first
differences stated in DESIGN.md:

* every LoRA matrix lives in ONE bf16 buffer and its gradient in ONE fp32 buffer (`FlatLoRA`); the lora_grad kernels
  accumulate straight into that buffer, data parallelism is a single NCCL all-reduce of it (the frozen NF4 base is
  replicated and never communicated), and clipping + AdamW + zero_grad are two kernels over it;
* the whole step (noise, forward, loss, backward, all-reduce, update) is captured once per (H, W) bucket in a CUDA graph
  and replayed: no host synchronisation, no `.item()`, no per-step tensor-map encoding.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops
from .jit import Denoiser, DenoiserConfig, JiT_B_16_Config, JiT_H_16_Config, JiT_L_16_Config
from .modules.peft import LoRAConfig, LoRALinear, PeftTargetConfig
from .modules.quant import quantize_inplace
from .modules.state_dict import RegexMatch

MODEL_CONFIGS = {"JiT-B/16": JiT_B_16_Config, "JiT-L/16": JiT_L_16_Config, "JiT-H/16": JiT_H_16_Config}
BLOCK_LINEARS = ("to_q", "to_k", "to_v", "to_o", "w_1", "w_2", "w_3")
LORA_TARGET = RegexMatch(regex=r"blocks\.\d+\.(attn|mlp)\.")   # see SURVEY 8b: ".mlp." alone also hits time_embedder.mlp


class ClassEncoder(nn.Module):
    """Label-id context encoder, src/models/jit/class_encoder.py:94-140: an embedding with a padding slot; the mask marks
    the leading valid labels of each sample."""

    def __init__(self, num_classes: int, embedding_dim: int):
        super().__init__()
        self.num_classes = num_classes
        self.pad_token_id = num_classes
        self.embedding = nn.Embedding(num_classes + 1, embedding_dim, padding_idx=num_classes)

    def forward(self, class_ids: torch.Tensor) -> torch.Tensor:
        return self.embedding(class_ids)


def build_jit_qlora(model: str | DenoiserConfig = "JiT-B/16", rank: int = 16, alpha: float = 16.0, device="cuda",
                    seed: int = 42, quantize: bool = True, lora_up_std: float = 0.0) -> Denoiser:
    """Random-init JiT (JiT.initialize_weights, reference denoiser.py:764-798) in bf16, block linears NF4-quantised
    (quantize_inplace -> bnb_nf4) and wrapped with LoRA (PeftTargetConfig.replace_to_peft_layer); only LoRA trains."""
    cfg = MODEL_CONFIGS[model]() if isinstance(model, str) else model
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = Denoiser(cfg)
    net.initialize_weights()
    net.to(torch.bfloat16)
    net.requires_grad_(False)
    if quantize:
        names = [n for n, _ in net.named_modules()]
        keys = [n for n in names if n.startswith("blocks.") and n.rsplit(".", 1)[-1] in BLOCK_LINEARS]
        quantize_inplace(net, "bnb_nf4", keys)
    PeftTargetConfig(include_keys=[LORA_TARGET], config=LoRAConfig(rank=rank, alpha=alpha)).replace_to_peft_layer(net)
    net.to(device)
    for name, p in net.named_parameters():
        if "lora_up" in name and lora_up_std > 0:
            nn.init.normal_(p, std=lora_up_std)
        p.requires_grad_(("lora_down" in name or "lora_up" in name) and "alpha" not in name)
    torch.random.set_rng_state(gen_state)
    return net


class FlatLoRA:
    """All trainable LoRA matrices of a model as views of one bf16 buffer, gradients as views of one fp32 buffer.

    `param._vpt_grad32` is the hook the kernels look for (ops.grad_sink): gradients are accumulated there by the
    lora_grad kernels and never pass through autograd's AccumulateGrad."""

    def __init__(self, model: nn.Module):
        self.params: list[nn.Parameter] = []
        mods = [(n, m) for n, m in model.named_modules() if isinstance(m, LoRALinear)]
        i = 0
        while i < len(mods):
            # q / k / v of one attention: the three lora_down matrices back to back, then the three lora_up matrices, so that
            # the fused block's single q | k | v GEMM takes them as two zero-copy views (ops.stacked)
            group = [mods[i]]
            stem = mods[i][0].rsplit(".", 1)[0]
            if mods[i][0].endswith(".to_q") and i + 2 < len(mods) and mods[i + 1][0] == f"{stem}.to_k" and mods[i + 2][0] == f"{stem}.to_v":
                group = mods[i:i + 3]
            for pick in ("lora_down", "lora_up") if len(group) == 3 else (None,):
                for _, mod in group:
                    for p in ((mod.lora_down.weight, mod.lora_up.weight) if pick is None else (getattr(mod, pick).weight,)):
                        if p.requires_grad:
                            self.params.append(p)
            i += len(group)
        if not self.params:
            raise ValueError("no trainable LoRA parameters")
        dev = self.params[0].device
        # every slice starts on a 16-byte boundary of the fp32 buffer (float4 loads) -> sizes rounded up to 8 elements
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 7) // 8 * 8
        self.numel = n
        self.param = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.direct = True
        for p, off in zip(self.params, self.offsets):
            view = self.param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            if p.dtype != torch.bfloat16:
                raise TypeError("LoRA matrices are bf16 on this path (PeftConfigMixin.dtype)")
            p._vpt_grad32 = self.grad[off:off + p.numel()].view_as(p)
            rank = min(p.shape)
            self.direct = self.direct and rank == ops.RANK

    def all_reduce(self, group=None) -> float:
        """The one exchange step of data parallelism (reference: DDP's bucketed all-reduce of the trainable = LoRA
        gradients, src/trainer/common.py:62-65,376-380): SUM all-reduce of the flat fp32 gradient buffer.  Returns the
        factor that turns the sum into the mean; the optimiser kernel folds it in (grad_scale)."""
        if group is None and not torch.distributed.is_initialized():
            return 1.0
        world = torch.distributed.get_world_size(group)
        if world > 1:
            torch.distributed.all_reduce(self.grad, group=group)
        return 1.0 / world

    def gather_autograd_grads(self) -> None:
        """Ranks other than 16 take the generic autograd route (`p.grad`); fold those into the flat buffer."""
        for p in self.params:
            if p.grad is not None:
                p._vpt_grad32.add_(p.grad.float())
                p.grad = None


@dataclass
class TrainHParams:
    lr: float = 1e-4
    betas: tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 0.01
    clip_grad_norm: float | None = 1.0
    loss_target: str = "image"        # configs/jit/x-loss/config.yml:21 ("image" | "velocity")
    timestep_eps: float = 0.05
    noise_scale: float = 1.0
    ts_std: float = 0.8               # scale_shift_sigmoid timestep sampling (src/modules/timestep/sampling.py:259-272)
    ts_mean: float = -0.8
    # "adamw" (torch.optim.AdamW semantics) or "radam_schedulefree" (schedulefree.RAdamScheduleFree, the optimiser of
    # the shipped YAMLs, configs/jit/x-loss/config.yml:75 with lr 1e-4; package defaults otherwise: no weight decay)
    optimizer: str = "adamw"
    # micro-steps per optimiser step; gradients are exchanged only on the last one (accelerator.no_sync +
    # gradient_accumulation_steps of the reference, src/trainer/common.py:322-331) and averaged over all of them
    grad_accum_steps: int = 1
    sf_r: float = 0.0
    sf_weight_lr_power: float = 2.0
    sf_silent_sgd_phase: bool = True


class TrainState:
    """What persists across steps and is shared by every (H, W) bucket of a run: the flat LoRA parameter / gradient
    buffers, the AdamW moments and step counter, the frozen class-label table, and the CUDA-graph memory pool (only one
    bucket's graph runs at a time, so all of them replay out of the same pool)."""

    def __init__(self, model: Denoiser, num_classes: int = 1000):
        dev = next(model.parameters()).device
        cfg = model.config
        gen = torch.Generator(device="cpu").manual_seed(1234)   # same class table on every rank (replicated, frozen)
        self.class_encoder = ClassEncoder(num_classes, cfg.context_dim)
        with torch.no_grad():
            self.class_encoder.embedding.weight.copy_(torch.randn(num_classes + 1, cfg.context_dim, generator=gen) * 0.02)
            self.class_encoder.embedding.weight[num_classes].zero_()
        self.class_encoder.to(dev, torch.bfloat16).requires_grad_(False)
        self.flat = FlatLoRA(model)
        n = self.flat.numel
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        # schedule-free optimiser state (allocated on first use): base sequence z, schedule scalars, coefficient scratch
        self.z: torch.Tensor | None = None
        self.sched = torch.zeros(4, dtype=torch.float64, device=dev)
        self.coef = torch.zeros(8, dtype=torch.float32, device=dev)
        self.eval_mode = False
        self.pool = None

    def tensors(self) -> tuple[torch.Tensor, ...]:
        extra = () if self.z is None else (self.z,)
        return (self.flat.param, self.exp_avg, self.exp_avg_sq, self.step_t, self.sched) + extra

    def ensure_z(self) -> torch.Tensor:
        if self.z is None:
            self.z = self.flat.param.float()
        return self.z

    def set_eval(self, eval_mode: bool, beta1: float) -> None:
        """optimizer.eval() / optimizer.train() of a schedule-free optimiser: the parameters become the averaged iterate x
        (what is evaluated and saved) or go back to the training sequence y.  No-op for AdamW (z is None)."""
        if self.z is None or eval_mode == self.eval_mode:
            return
        ops.radam_schedulefree_swap(self.flat.param, self.z, beta1, eval_mode)
        self.eval_mode = eval_mode


def init_communicator(group, device) -> None:
    """Communicator set-up where every rank is in lock-step (construction of the trainer), never inside a lazily
    triggered warm-up: one 4-byte all-reduce."""
    if group is None or torch.distributed.get_world_size(group) == 1:
        return
    torch.distributed.all_reduce(torch.zeros(1, device=device), group=group)
    torch.cuda.synchronize(device)


class JiTQLoRATrainStep:
    """One optimisation step of JiT NF4-QLoRA class-to-image training; `run()` replays a captured CUDA graph.

    Inputs of a step (static device buffers the caller fills, e.g. by an async copy from pinned host memory):
      image [B,3,H,W] fp16 (the dataset emits fp16, src/dataset/text_to_image.py:152), class_ids [B,T] int64,
      attention_mask [B,T] int64 (leading ones).  Output: `loss` (fp32 device scalar of the last step).

    Data parallelism (world > 1): for a step captured in lock-step on every rank (`capture(in_lockstep=True)`,
    JiTQLoRATrainer.precapture) the LoRA-gradient all-reduce is part of the captured graph (`nccl_in_graph`: a step is ONE
    graph launch); a step captured lazily by one rank keeps it between two graphs.  `overlap_chunks` > 1 splits the
    exchange so that the gradient ranges of the last blocks -- final long before backward ends -- are reduced on a side
    stream under the remaining backward (measured slower than one exchange at the end: the default is 1).  No collective
    is ever issued by a warm-up or a capture: ranks may capture new (H, W) buckets at different times without their
    collectives falling out of step."""

    def __init__(self, model: Denoiser, batch: int, height: int, width: int, num_classes: int = 1000,
                 max_token_length: int = 64, hp: TrainHParams | None = None, process_group=None, use_graph: bool = True,
                 seed: int | None = 0, state: TrainState | None = None, nccl_in_graph: bool | None = None,
                 overlap_chunks: int | None = None):
        self.model = model
        self.hp = hp or TrainHParams()
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        dev = next(model.parameters()).device
        self.device = dev
        cfg = model.config
        fresh_state = state is None
        self.state = state if state is not None else TrainState(model, num_classes)
        self.class_encoder = self.state.class_encoder
        self.flat = self.state.flat
        self.exp_avg, self.exp_avg_sq = self.state.exp_avg, self.state.exp_avg_sq
        self.step_t, self.sumsq = self.state.step_t, self.state.sumsq
        self.image = torch.zeros((batch, cfg.in_channels, height, width), dtype=torch.float16, device=dev)
        self.class_ids = torch.full((batch, max_token_length), num_classes, dtype=torch.int64, device=dev)
        self.attention_mask = torch.zeros((batch, max_token_length), dtype=torch.int64, device=dev)
        self.size_info = torch.tensor([[height, width]], device=dev).repeat(batch, 1)
        self.crop = torch.zeros_like(self.size_info)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.use_graph = use_graph
        self.graph: torch.cuda.CUDAGraph | None = None          # the whole step (or compute only, see capture())
        self.graph_update: torch.cuda.CUDAGraph | None = None   # clip + update, when the exchange stays outside the graphs
        self.graph_micro: torch.cuda.CUDAGraph | None = None    # a gradient-accumulation micro-step (compute only)
        self.kernel_launches = 0
        nccl = self.world > 1 and torch.distributed.get_backend(process_group) == "nccl"
        if nccl_in_graph is None:
            nccl_in_graph = os.environ.get("VPT_NCCL_IN_GRAPH", "1") != "0"
        self.nccl_in_graph = bool(nccl_in_graph) and nccl
        if overlap_chunks is None:
            # 1 = one all-reduce after backward.  Measured on 2 B200 (profiles/r2f_dp_variants.txt): splitting it so that the
            # last blocks' chunk runs under the rest of backward is SLOWER (14.62 vs 14.46 ms / step, JiT-B): the NCCL kernel
            # takes SMs from the persistent GEMMs for longer than the 11 MB exchange costs when it runs alone
            overlap_chunks = int(os.environ.get("VPT_DP_CHUNKS", "1"))
        self._chunks = self._chunk_plan(max(1, overlap_chunks)) if self.world > 1 else []
        self._side = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self._sent = 0                   # chunks of this step whose all-reduce has been issued
        self._collectives = False        # only a real step (never a warm-up) may issue collectives
        if fresh_state:
            init_communicator(process_group, dev)
        if seed is not None:
            torch.manual_seed(seed)
        model.train()

    # ------------------------------------------------------------------ gradient exchange
    def _chunk_plan(self, n_chunks: int) -> list[tuple[int, int, int]]:
        """(first_block, lo, hi): the flat-gradient range [lo, hi) belongs to blocks >= first_block and is complete once
        backward has passed `first_block`.  Needs the flat buffer to be ordered by block (it is: FlatLoRA walks the
        modules in order) with every trainable matrix inside a block; otherwise one chunk after backward."""
        names = {id(p): n for n, p in self.model.named_parameters()}
        depth = len(self.model.blocks)
        start = [None] * depth
        ok = True
        for p, off in zip(self.flat.params, self.flat.offsets):
            n = names.get(id(p), "")
            if not n.startswith("blocks."):
                ok = False
                break
            b = int(n.split(".")[1])
            start[b] = off if start[b] is None else min(start[b], off)
        ok = ok and all(st is not None for st in start) and all(start[i] < start[i + 1] for i in range(depth - 1))
        total = self.flat.numel
        if not ok or n_chunks <= 1 or depth < 2 * n_chunks:
            return [(0, 0, total)]
        plan, hi = [], total
        for c in range(n_chunks - 1, 0, -1):          # last blocks first: they finish backward first
            fb = depth * c // n_chunks
            plan.append((fb, start[fb], hi))
            hi = start[fb]
        plan.append((0, 0, hi))
        return plan

    def _send_chunk(self, lo: int, hi: int) -> None:
        main = torch.cuda.current_stream(self.device)
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            torch.distributed.all_reduce(self.flat.grad[lo:hi], group=self.group)

    def _after_block_backward(self, block_index: int) -> None:
        """Called by JiTBlockFn.backward once block `block_index` has accumulated its LoRA gradients."""
        if not self._collectives:
            return
        while self._sent < len(self._chunks) - 1 and self._chunks[self._sent][0] >= block_index:
            _, lo, hi = self._chunks[self._sent]
            self._send_chunk(lo, hi)
            self._sent += 1

    def _exchange(self) -> float:
        """The exchange step of data parallelism: SUM all-reduce of the flat LoRA-gradient buffer (NCCL), in chunks of
        which all but the last were already issued during backward; returns 1 / world."""
        if self.world == 1:
            return 1.0
        if self._collectives:
            for _, lo, hi in self._chunks[self._sent:]:
                self._send_chunk(lo, hi)
            self._sent = 0
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        return 1.0 / self.world

    # ------------------------------------------------------------------ the step itself (eager or under capture)
    def _compute(self) -> None:
        """Noise, forward, loss, backward: LoRA gradients end up in the flat fp32 buffer."""
        hp = self.hp
        images = self.image
        B = images.shape[0]
        with torch.no_grad():
            context = self.class_encoder(self.class_ids)
            t = (torch.randn(B, device=self.device) * hp.ts_std + hp.ts_mean).sigmoid()
            # prepare_scaled_noised_latents (reference flow_match.py:60-74) as one kernel, same roundings
            if ops.FUSED_GLUE:
                noisy, noisy_bf16 = ops.noise_mix(images, torch.randn_like(images), t, hp.noise_scale)
            else:
                noise = torch.randn_like(images) * hp.noise_scale
                tv = t.view(B, 1, 1, 1).to(images.dtype)
                noisy = tv * images + (1 - tv) * noise
                noisy_bf16 = noisy.to(torch.bfloat16)
        pred = self.model(image=noisy_bf16, timestep=t.to(torch.bfloat16), context=context,
                          original_size=self.size_info, target_size=self.size_info, crop_coords=self.crop,
                          context_mask=self.attention_mask)
        loss = ops.flow_loss(pred, images, noisy, t, loss_target=hp.loss_target, clamp_eps=hp.timestep_eps)
        self._sent = 0
        from .jit import denoiser as _dn
        _dn.BLOCK_BACKWARD_HOOK = self._after_block_backward if self.world > 1 else None
        try:
            loss.backward()
        finally:
            _dn.BLOCK_BACKWARD_HOOK = None
        if not self.flat.direct:
            self.flat.gather_autograd_grads()
        self.loss.copy_(loss.detach())

    def _update(self, scale: float) -> None:
        """Gradient-norm clipping + AdamW + zero_grad over the flat buffers (two kernels)."""
        hp = self.hp
        scale = scale / max(1, hp.grad_accum_steps)
        sumsq = None
        if hp.clip_grad_norm is not None:
            self.sumsq.zero_()
            ops.grad_sumsq(self.flat.grad, scale, self.sumsq)
            sumsq = self.sumsq
        self.step_t += 1
        if hp.optimizer == "radam_schedulefree":
            ops.radam_schedulefree_step(self.flat.param, self.flat.grad, self.state.ensure_z(), self.exp_avg_sq, self.state.sched,
                                        self.state.coef, hp.lr, hp.betas, hp.eps, hp.weight_decay, hp.sf_r, hp.sf_weight_lr_power,
                                        hp.sf_silent_sgd_phase, grad_scale=scale, sumsq=sumsq,
                                        max_norm=hp.clip_grad_norm or 0.0, zero_grad=True)
            return
        if hp.optimizer != "adamw":
            raise ValueError(f"unknown optimizer '{hp.optimizer}'")
        ops.adamw_step(self.flat.param, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.step_t, hp.lr, hp.betas, hp.eps,
                       hp.weight_decay, grad_scale=scale, sumsq=sumsq, max_norm=hp.clip_grad_norm or 0.0, zero_grad=True)

    def _step(self, sync: bool = True) -> None:
        self._compute()
        if sync:
            self._update(self._exchange())

    def _warmup(self, n: int) -> None:
        """Eager steps WITHOUT collectives on a side stream, with every piece of training state restored afterwards: a
        new (H, W) bucket may be captured in the middle of a run, by one rank only."""
        snap = [t.clone() for t in self.state.tensors()]
        grad = self.flat.grad.clone()                            # a capture may fall between accumulation micro-steps
        rng = torch.cuda.get_rng_state(self.device)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(n):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, t0 in zip(self.state.tensors(), snap):
            t.copy_(t0)
        self.flat.grad.copy_(grad)
        torch.cuda.set_rng_state(rng, self.device)
        torch.cuda.synchronize()

    def _graph_ctx(self, g: torch.cuda.CUDAGraph):
        kw = {} if self.state.pool is None else {"pool": self.state.pool}
        if self.world > 1:
            kw["capture_error_mode"] = "thread_local"            # NCCL's watchdog thread may touch CUDA during a capture
        return torch.cuda.graph(g, **kw)

    def capture(self, warmup: int = 2, in_lockstep: bool = False) -> None:
        """Eager warm-up (one-time kernel attribute setup, allocator growth), then capture.

        `in_lockstep=True` (JiTQLoRATrainer.precapture: every rank captures the same sequence of buckets at the same time):
        at world > 1 the chunked NCCL all-reduce is captured INTO the step's graph (fork / join of the side stream become
        graph edges), so a step is one graph launch.  A capture that only this rank performs (a bucket met for the first
        time in the middle of a run) keeps the collective OUT of its graphs -- compute | update with the all-reduce
        launched between them: NCCL may exchange buffer registrations between ranks while a collective is being captured,
        which a lone rank would wait for forever."""
        if self.hp.optimizer == "radam_schedulefree":
            self.state.ensure_z()
        self._collectives = False
        self._warmup(warmup)
        snap = [t.clone() for t in self.state.tensors()]         # capturing does not execute, but be safe about state
        grad = self.flat.grad.clone()
        rng = torch.cuda.get_rng_state(self.device)
        before = ops._lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        if self.world == 1 or (self.nccl_in_graph and in_lockstep):
            self._collectives = self.world > 1
            with self._graph_ctx(self.graph):
                self._step()
            self.state.pool = self.graph.pool()
        else:
            with self._graph_ctx(self.graph):
                self._compute()
            self.state.pool = self.graph.pool()
            self.graph_update = torch.cuda.CUDAGraph()
            with self._graph_ctx(self.graph_update):
                self._update(1.0 / self.world)
        self._collectives = False
        self.kernel_launches = ops._lib.launch_count() - before
        for t, t0 in zip(self.state.tensors(), snap):
            t.copy_(t0)
        self.flat.grad.copy_(grad)
        torch.cuda.set_rng_state(rng, self.device)
        torch.cuda.synchronize()

    def capture_micro(self) -> None:
        """Graph of a gradient-accumulation micro-step: compute only, gradients keep accumulating in the flat buffer."""
        if self.graph is None:
            self.capture()
        snap = [t.clone() for t in self.state.tensors()]
        grad = self.flat.grad.clone()
        rng = torch.cuda.get_rng_state(self.device)
        self.graph_micro = torch.cuda.CUDAGraph()
        with self._graph_ctx(self.graph_micro):
            self._compute()
        for t, t0 in zip(self.state.tensors(), snap):
            t.copy_(t0)
        self.flat.grad.copy_(grad)
        torch.cuda.set_rng_state(rng, self.device)
        torch.cuda.synchronize()

    def run(self, sync: bool = True) -> torch.Tensor:
        """sync=False: a micro-step of gradient accumulation -- no exchange, no update (`no_sync` of the reference)."""
        if self.use_graph:
            if self.graph is None:
                self.capture()
            if not sync:
                if self.graph_micro is None:
                    self.capture_micro()
                self.graph_micro.replay()
            elif self.graph_update is None:
                self.graph.replay()
            else:
                self.graph.replay()
                self._collectives = True
                self._exchange()
                self._collectives = False
                self.graph_update.replay()
        else:
            before = ops._lib.launch_count()
            self._collectives = self.world > 1 and sync
            try:
                self._step(sync)
            finally:
                self._collectives = False
            self.kernel_launches = ops._lib.launch_count() - before
        return self.loss


class _Staged:
    """One staging slot of the input pipeline: device copies of a host batch made on the copy stream."""
    __slots__ = ("image", "class_ids", "attention_mask", "ready", "free", "key")

    def __init__(self, step: "JiTQLoRATrainStep"):
        self.image = torch.empty_like(step.image)
        self.class_ids = torch.empty_like(step.class_ids)
        self.attention_mask = torch.empty_like(step.attention_mask)
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self.key = None


class JiTQLoRATrainer:
    """The caller-facing loop body for aspect-ratio-bucketed training (BASELINE.json configs[3]: (H, W) differs per step
    and per rank; reference: src/dataset/aspect_ratio_bucket.py:20-60 feeding train/jit/class_to_image.py:166-243).

    One `JiTQLoRATrainStep` (= one CUDA graph) per (batch, H, W) bucket, created on first use; all of them share the
    LoRA parameters, gradients, AdamW state and graph memory pool of one `TrainState`.  `train_step` takes the host
    batch exactly as the reference's dataloader yields it (image fp16 [B,3,H,W], class ids, mask) and replays the graph.

    Input pipeline: `train_step(batch, prefetch=next_batch)` starts the NEXT batch's host-to-device copy on a copy stream
    into one of two staging slots of its bucket while this step computes; the step then only pays a device-to-device copy
    of its own (already resident) batch into the graph's input buffers.  The graph itself is untouched, so a prefetched
    run is bit-identical to a plain one.  The loss is returned as a device scalar (no sync); `read_loss()` gives the
    loss of a finished earlier step from pinned host memory without stalling the current one.

    Checkpoints: the adapter as safetensors with the reference's key names (`get_adapter_parameters`,
    src/modules/peft/functional.py:114-125: `<path>.lora_down.weight`, `<path>.lora_up.weight`, `<path>.alpha`) and the
    optimiser state (moments by the same names + the step counter) beside it, so a run resumes bit-exactly."""

    def __init__(self, model: Denoiser, num_classes: int = 1000, max_token_length: int = 64,
                 hp: TrainHParams | None = None, process_group=None, use_graph: bool = True, seed: int = 0,
                 nccl_in_graph: bool | None = None, overlap_chunks: int | None = None):
        self.model = model
        self.num_classes, self.max_token_length = num_classes, max_token_length
        self.hp = hp or TrainHParams()
        self.group = process_group
        self.use_graph = use_graph
        self.nccl_in_graph, self.overlap_chunks = nccl_in_graph, overlap_chunks
        self.state = TrainState(model, num_classes)
        self.device = next(model.parameters()).device
        init_communicator(process_group, self.device)
        self.buckets: dict[tuple[int, int, int], JiTQLoRATrainStep] = {}
        self._staging: dict[tuple[int, int, int], list[_Staged]] = {}
        self._stage_turn = 0
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._micro = 0
        # losses of the last steps in pinned host memory (ring), each guarded by an event
        self._loss_ring = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._loss_events = [torch.cuda.Event() for _ in range(8)]
        self._steps_run = 0
        torch.manual_seed(seed)

    def bucket(self, batch: int, height: int, width: int) -> JiTQLoRATrainStep:
        key = (batch, height, width)
        step = self.buckets.get(key)
        if step is None:
            step = JiTQLoRATrainStep(self.model, batch, height, width, self.num_classes, self.max_token_length, self.hp,
                                     self.group, self.use_graph, seed=None, state=self.state,
                                     nccl_in_graph=self.nccl_in_graph, overlap_chunks=self.overlap_chunks)
            self.buckets[key] = step
        return step

    def precapture(self, shapes) -> None:
        """Capture the graphs of the given (batch, H, W) buckets up front.  COLLECTIVE at world > 1: every rank must call it
        with the same list at the same point -- in exchange the gradient all-reduce is captured into each step's graph
        (one launch per step, chunks overlapped with backward).  Buckets that are not precaptured are captured lazily by
        the rank that meets them, with the all-reduce between two graphs; both kinds interoperate (the collective is the
        same), so ranks may meet new buckets at different times."""
        for b, h, w in shapes:
            step = self.bucket(b, h, w)
            if step.use_graph and step.graph is None:
                step.capture(in_lockstep=True)

    def close(self) -> None:
        """Drop every captured graph (and the shared graph pool).  Call before torch.distributed.destroy_process_group():
        tearing down a NCCL communicator whose collectives are still referenced by live CUDA graphs blocks in
        ncclCommDestroy (observed with NCCL 2.28.9)."""
        import gc
        torch.cuda.synchronize(self.device)
        for step in self.buckets.values():
            step.graph = step.graph_update = step.graph_micro = None
        self.buckets.clear()
        self._staging.clear()
        self.state.pool = None
        gc.collect()
        torch.cuda.synchronize(self.device)

    @staticmethod
    def _check_host_mask(attention_mask: torch.Tensor) -> None:
        if attention_mask.is_cuda:
            return                                   # device masks are checked by prefix_key_lengths (device-side assert)
        from .modules.attention import prefix_key_lengths
        prefix_key_lengths(attention_mask)           # raises for float / non-prefix masks

    def prefetch(self, image: torch.Tensor, class_ids: torch.Tensor, attention_mask: torch.Tensor) -> None:
        """Start the host-to-device copy of a FUTURE step's batch on the copy stream (pinned host tensors)."""
        self._check_host_mask(attention_mask)
        B, _, H, W = image.shape
        key = (B, H, W)
        step = self.bucket(B, H, W)
        slots = self._staging.setdefault(key, [])
        slot = next((sl for sl in slots if sl.key is None), None)       # a slot whose batch has been consumed
        if slot is None:
            if len(slots) < 2:
                slots.append(_Staged(step))
                slot = slots[-1]
            else:                                                        # more than two batches ahead: replace the older
                slot = slots[self._stage_turn % 2]
                self._stage_turn += 1
        cs = self._copy_stream
        cs.wait_event(slot.free)                     # the step that last read this slot has copied it out
        with torch.cuda.stream(cs):
            slot.image.copy_(image, non_blocking=True)
            slot.class_ids.copy_(class_ids, non_blocking=True)
            slot.attention_mask.copy_(attention_mask, non_blocking=True)
            slot.ready.record(cs)
        slot.key = (id(image), id(class_ids), id(attention_mask))

    def _staged_for(self, key, image, class_ids, attention_mask) -> _Staged | None:
        want = (id(image), id(class_ids), id(attention_mask))
        for slot in self._staging.get(key, ()):
            if slot.key == want:
                return slot
        return None

    def train_step(self, image: torch.Tensor, class_ids: torch.Tensor, attention_mask: torch.Tensor,
                   prefetch: tuple | None = None) -> torch.Tensor:
        B, _, H, W = image.shape
        if self.state.eval_mode:
            raise RuntimeError("train_step in eval mode: call trainer.train() first (schedule-free optimiser)")
        step = self.bucket(B, H, W)
        if step.use_graph and step.graph is None:
            step.capture()                       # before the copies: warm-up must not consume this batch's buffers
        main = torch.cuda.current_stream(self.device)
        slot = self._staged_for((B, H, W), image, class_ids, attention_mask)
        if slot is not None:
            main.wait_event(slot.ready)
            step.image.copy_(slot.image, non_blocking=True)
            step.class_ids.copy_(slot.class_ids, non_blocking=True)
            step.attention_mask.copy_(slot.attention_mask, non_blocking=True)
            slot.free.record(main)
            slot.key = None
        else:
            self._check_host_mask(attention_mask)
            step.image.copy_(image, non_blocking=True)
            step.class_ids.copy_(class_ids, non_blocking=True)
            step.attention_mask.copy_(attention_mask, non_blocking=True)
        if prefetch is not None:
            self.prefetch(*prefetch)
        accum = max(1, self.hp.grad_accum_steps)
        self._micro += 1
        sync = self._micro % accum == 0
        loss = step.run(sync=sync)
        i = self._steps_run % 8
        self._loss_ring[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_events[i].record(main)
        self._steps_run += 1
        return loss

    def read_loss(self, steps_back: int = 1) -> float:
        """Loss of the step `steps_back` calls ago (1 = the previous call's ... 0 = the call just made, which waits for
        it), read from pinned host memory after its own event: reading an older step never stalls the running one."""
        n = self._steps_run - 1 - steps_back
        if n < 0 or steps_back >= 8:
            raise IndexError("no such step in the loss ring")
        i = n % 8
        self._loss_events[i].synchronize()
        return float(self._loss_ring[i])

    @property
    def global_step(self) -> int:
        return int(self.state.step_t.item())

    @property
    def scheduled_lr(self) -> float:
        """param_group["scheduled_lr"] of a schedule-free optimiser (what src/trainer/common.py:499-506 logs)."""
        return float(self.state.sched[3].item()) if self.hp.optimizer == "radam_schedulefree" else self.hp.lr

    def eval(self) -> None:
        """Before validation / saving with a schedule-free optimiser (optimizer.eval()): parameters -> averaged iterate."""
        self.state.set_eval(True, self.hp.betas[0])

    def train(self) -> None:
        self.state.set_eval(False, self.hp.betas[0])

    # ------------------------------------------------------------------ checkpoint / resume
    def _named_slices(self):
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p, off in zip(self.state.flat.params, self.state.flat.offsets):
            yield names[id(p)], off, p

    def adapter_state_dict(self) -> dict[str, torch.Tensor]:
        from .modules.peft import get_adapter_parameters
        return {k: v.detach().cpu().contiguous() for k, v in get_adapter_parameters(self.model).items()}

    def save_checkpoint(self, directory: str) -> None:
        import os

        from safetensors.torch import save_file
        os.makedirs(directory, exist_ok=True)
        save_file(self.adapter_state_dict(), os.path.join(directory, "adapter.safetensors"),
                  metadata={"format": "pt", "peft": "lora"})
        opt = {"step": self.state.step_t.detach().cpu().clone(), "sched": self.state.sched.detach().cpu().clone()}
        for name, off, p in self._named_slices():
            n = p.numel()
            opt[f"{name}.exp_avg"] = self.state.exp_avg[off:off + n].view_as(p).cpu().clone()
            opt[f"{name}.exp_avg_sq"] = self.state.exp_avg_sq[off:off + n].view_as(p).cpu().clone()
            if self.state.z is not None:
                opt[f"{name}.z"] = self.state.z[off:off + n].view_as(p).cpu().clone()
        save_file(opt, os.path.join(directory, "optimizer.safetensors"))
        torch.save({"rng_cpu": torch.random.get_rng_state(), "rng_cuda": torch.cuda.get_rng_state(), "hp": vars(self.hp)},
                   os.path.join(directory, "trainer_state.pt"))

    def load_checkpoint(self, directory: str, strict: bool = True) -> None:
        import os

        from safetensors.torch import load_file
        adapter = load_file(os.path.join(directory, "adapter.safetensors"))
        opt_path = os.path.join(directory, "optimizer.safetensors")
        opt = load_file(opt_path) if os.path.exists(opt_path) else None
        with torch.no_grad():
            for name, off, p in self._named_slices():
                if name not in adapter:
                    if strict:
                        raise KeyError(f"adapter checkpoint has no '{name}'")
                    continue
                p.copy_(adapter[name].to(p.device, p.dtype))          # in place: the parameter is a view of the flat buffer
                if opt is not None:
                    n = p.numel()
                    self.state.exp_avg[off:off + n].copy_(opt[f"{name}.exp_avg"].reshape(-1))
                    self.state.exp_avg_sq[off:off + n].copy_(opt[f"{name}.exp_avg_sq"].reshape(-1))
                    if f"{name}.z" in opt:
                        self.state.ensure_z()[off:off + n].copy_(opt[f"{name}.z"].reshape(-1))
            if opt is not None:
                self.state.step_t.copy_(opt["step"])
                if "sched" in opt:
                    self.state.sched.copy_(opt["sched"])
        ts = os.path.join(directory, "trainer_state.pt")
        if os.path.exists(ts):
            st = torch.load(ts, weights_only=False)
            torch.random.set_rng_state(st["rng_cpu"])
            torch.cuda.set_rng_state(st["rng_cuda"])


def synthetic_batch(batch: int, height: int, width: int, num_classes: int = 1000, max_token_length: int = 64,
                    seed: int = 0, pin: bool = True):
    """Host-side synthetic batch of the named shape: image ~ N(0,1) fp16, 8..40 valid labels per sample."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn((batch, 3, height, width), generator=g).to(torch.float16)
    n_labels = torch.randint(8, 41, (batch,), generator=g)
    ids = torch.randint(0, num_classes, (batch, max_token_length), generator=g)
    ar = torch.arange(max_token_length).unsqueeze(0)
    mask = (ar < n_labels.unsqueeze(1)).to(torch.int64)
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, num_classes))
    out = (image, ids, mask)
    if pin and torch.cuda.is_available():
        out = tuple(t.pin_memory() for t in out)
    return out


def step_flops(cfg: DenoiserConfig, batch: int, height: int, width: int, ctx_len: int = 64, rank: int = 16) -> dict:
    """Algorithmic FLOPs of one training step (fwd + bwd, blocks only): per linear 4 M K N + 6 M r (K+N),
    attention 4 B H L^2 hd forward and 2.5x that backward (full L x L counted)."""
    D, H = cfg.hidden_size, cfg.num_heads
    hd = D // H
    F = int(int(D * cfg.mlp_ratio) * 2 / 3)
    n_patch = (height // cfg.patch_size) * (width // cfg.patch_size)
    pre = n_patch + 6 + cfg.num_time_tokens
    lin = attn = 0.0
    for i in range(cfg.depth):
        L = pre + (ctx_len if i >= cfg.context_start_block else 0)
        M = batch * L
        for (K, N) in ((D, D),) * 4 + ((D, F),) * 2 + ((F, D),):
            lin += 4.0 * M * K * N + 6.0 * M * rank * (K + N)
        attn += 3.5 * 4.0 * batch * H * L * L * hd
    return {"linear": lin, "attention": attn, "total": lin + attn}
