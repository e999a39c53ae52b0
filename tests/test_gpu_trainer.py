"""Bucketed trainer (one CUDA graph per (H, W) bucket over shared LoRA / optimiser state), checkpoint / resume, and the
NF4 prequantisation tool: the callers and data formats either side of the hot path."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_cfg():
    from vision_pt_b200.jit import DenoiserConfig
    return DenoiserConfig(patch_size=16, in_channels=3, out_channels=3, hidden_size=128, depth=2, num_heads=2, mlp_ratio=4.0,
                          bottleneck_dim=32, num_time_tokens=4, rope_axes_dims=[16, 24, 24], context_dim=64,
                          context_start_block=1)


def _trainer(use_graph, seed=5):
    from vision_pt_b200 import train as T
    net = T.build_jit_qlora(_small_cfg(), rank=16, alpha=16.0, device="cuda", seed=11, lora_up_std=0.02)
    hp = T.TrainHParams(lr=2e-3, clip_grad_norm=1.0)
    return T, T.JiTQLoRATrainer(net, num_classes=10, max_token_length=16, hp=hp, use_graph=use_graph, seed=seed)


@pytest.mark.parametrize("use_graph", [True, False])
def test_buckets_share_one_training_state(use_graph):
    T, tr = _trainer(use_graph)
    shapes = [(8, 64, 64), (8, 64, 128), (8, 128, 64)]
    batches = [T.synthetic_batch(B, H, W, num_classes=10, max_token_length=16, seed=i) for i, (B, H, W) in enumerate(shapes)]
    p0 = tr.state.flat.param.clone()
    losses = []
    for it in range(9):
        loss = tr.train_step(*batches[it % 3])
        assert tr.global_step == it + 1           # capturing a new bucket's graph (2 warm-up steps) does not train
        losses.append(float(loss))
    assert len(tr.buckets) == 3 and all(torch.isfinite(torch.tensor(losses)))
    assert not torch.equal(tr.state.flat.param, p0)
    assert float(tr.state.flat.grad.abs().max()) == 0.0      # zero_grad is part of the update kernel
    # every bucket works on the same buffers
    for step in tr.buckets.values():
        assert step.flat is tr.state.flat and step.exp_avg is tr.state.exp_avg and step.step_t is tr.state.step_t


@pytest.mark.parametrize("use_graph", [True, False])
def test_checkpoint_resume_is_bit_exact(tmp_path, use_graph):
    T, tr = _trainer(use_graph)
    a = T.synthetic_batch(8, 64, 64, num_classes=10, max_token_length=16, seed=1)
    b = T.synthetic_batch(8, 64, 128, num_classes=10, max_token_length=16, seed=2)
    for batch in (a, b, a):
        tr.train_step(*batch)
    torch.cuda.synchronize()
    tr.save_checkpoint(str(tmp_path))
    for batch in (b, a):
        tr.train_step(*batch)
    torch.cuda.synchronize()
    want = (tr.state.flat.param.clone(), tr.state.exp_avg.clone(), tr.state.exp_avg_sq.clone(), tr.global_step)

    from safetensors.torch import load_file
    adapter = load_file(os.path.join(str(tmp_path), "adapter.safetensors"))
    # the reference's adapter key names (tests/test_peft.py:111-127 of the reference)
    assert "blocks.0.attn.to_q.lora_down.weight" in adapter and "blocks.1.mlp.w_3.lora_up.weight" in adapter
    assert "blocks.0.attn.to_q.alpha" in adapter and not any(".linear." in k for k in adapter)

    T2, tr2 = _trainer(use_graph, seed=999)       # a fresh process would start like this: same base, untrained adapter
    tr2.load_checkpoint(str(tmp_path))
    assert tr2.global_step == 3
    for batch in (b, a):
        tr2.train_step(*batch)
    torch.cuda.synchronize()
    assert tr2.global_step == want[3]
    assert torch.equal(tr2.state.flat.param, want[0])
    assert torch.equal(tr2.state.exp_avg, want[1]) and torch.equal(tr2.state.exp_avg_sq, want[2])


def test_adapter_file_loads_through_the_reference_style_api(tmp_path):
    """adapter.safetensors -> load_peft_weight on a model without adapters (reference: src/modules/peft/functional.py
    `load_peft_weight`): the wrapped model computes the same output as the trained one."""
    from safetensors.torch import load_file

    from vision_pt_b200 import train as T
    from vision_pt_b200.modules.peft import load_peft_weight
    T, tr = _trainer(False)
    batch = T.synthetic_batch(8, 64, 64, num_classes=10, max_token_length=16, seed=1)
    for _ in range(2):
        tr.train_step(*batch)
    tr.save_checkpoint(str(tmp_path))
    base = T.build_jit_qlora(_small_cfg(), rank=16, alpha=16.0, device="cuda", seed=11, lora_up_std=0.0)
    load_peft_weight(base, {k: v.cuda() for k, v in load_file(os.path.join(str(tmp_path), "adapter.safetensors")).items()})
    x = torch.randn(2, 3, 64, 64, device="cuda", dtype=torch.bfloat16)
    t = torch.tensor([0.3, 0.7], device="cuda", dtype=torch.bfloat16)
    ctx = torch.randn(2, 16, 64, device="cuda", dtype=torch.bfloat16)
    mask = torch.ones(2, 16, dtype=torch.int64, device="cuda")
    size = torch.tensor([[64, 64]] * 2, device="cuda")
    kw = dict(timestep=t, context=ctx, original_size=size, target_size=size, crop_coords=torch.zeros_like(size), context_mask=mask)
    with torch.no_grad():
        y1 = tr.model.eval()(image=x, **kw)
        y2 = base.eval()(image=x, **kw)
    assert torch.equal(y1, y2)


def test_quantize_model_tool_roundtrip(tmp_path):
    """fp checkpoint -> tools/quantize_model.py -> prequantised safetensors with bitsandbytes' key set -> loads into NF4Linear
    modules through replace_by_prequantized_weights (reference flow: src/modules/quant/functional.py:332-371)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import quantize_model as qm
    from safetensors.torch import load_file, save_file

    from vision_pt_b200.modules.quant import NF4Linear, replace_by_prequantized_weights
    torch.manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = torch.nn.ModuleDict({"to_q": torch.nn.Linear(128, 128), "to_out": torch.nn.Linear(128, 64)})
            self.head = torch.nn.Linear(64, 10)

    net = Net().to(torch.bfloat16)
    src, dst = str(tmp_path / "m.safetensors"), str(tmp_path / "m.nf4.safetensors")
    save_file({k: v.contiguous() for k, v in net.state_dict().items()}, src)
    rep = qm.quantize_file(src, dst, include=[r"^attn\..*\.weight$"], exclude=[], verify=True)
    assert rep["quantized"] == 2 and rep["bytes_after"] < rep["bytes_before"] and rep["max_abs_err"] < 0.15
    sd = load_file(dst)
    for name in ("attn.to_q.weight", "attn.to_out.weight"):
        assert sd[name].dtype == torch.uint8
        for suffix in ("absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state.bitsandbytes__nf4"):
            assert f"{name}.{suffix}" in sd
    assert sd["head.weight"].dtype == torch.bfloat16
    fresh = Net().to(torch.bfloat16)
    replace_by_prequantized_weights(fresh, sd)
    assert isinstance(fresh.attn["to_q"], NF4Linear) and isinstance(fresh.head, torch.nn.Linear)
    fresh.load_state_dict(sd)
    fresh.cuda()
    x = torch.randn(64, 128, device="cuda", dtype=torch.bfloat16)
    y = fresh.attn["to_q"](x)
    w = fresh.attn["to_q"].dequantize().float()
    ref = x.float() @ w.t() + fresh.attn["to_q"].bias.float()
    assert (y.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    assert (w.cpu() - net.attn["to_q"].weight.float()).abs().max() <= 0.15 * net.attn["to_q"].weight.float().abs().max()


def test_radam_schedulefree_kernel_matches_oracle():
    """vpt_radam_schedulefree_step (device-side schedule scalars, bf16 y / fp32 z) against oracle/optim.py with y rounded
    to bf16 after every step, 40 steps with gradient clipping, through the silent phase into the Adam phase."""
    from oracle import optim as oo
    from vision_pt_b200 import ops
    torch.manual_seed(3)
    n, lr, wd, max_norm = 4096 + 8, 5e-3, 0.01, 1.0
    y = (torch.randn(n) * 0.1).to(torch.bfloat16)
    dev_y = y.clone().cuda()
    z = dev_y.float()
    v = torch.zeros(n, device="cuda")
    g32 = torch.zeros(n, device="cuda")
    sched = torch.zeros(4, dtype=torch.float64, device="cuda")
    coef = torch.zeros(8, device="cuda")
    sumsq = torch.zeros(1, device="cuda")
    gr = oo.RAdamSFGroup(lr=lr, weight_decay=wd)
    ry = [y.double()]
    for step in range(40):
        g = torch.randn(n) * (3.0 if step % 7 == 0 else 0.01)       # sometimes clipped, sometimes not
        g32.copy_(g)
        sumsq.zero_()
        ops.grad_sumsq(g32, 1.0, sumsq)
        ops.radam_schedulefree_step(dev_y, g32, z, v, sched, coef, lr, weight_decay=wd, sumsq=sumsq, max_norm=max_norm)
        clip = min(1.0, max_norm / (float(g.double().norm()) + 1e-6))
        oo.radam_schedulefree_step(gr, ry, [g.double() * clip])
        ry[0].copy_(ry[0].float().to(torch.bfloat16).double())
        assert float(g32.abs().max()) == 0.0
    assert abs(float(sched[0]) - 40) == 0 and abs(float(sched[3]) - gr.scheduled_lr) <= 1e-9 * lr
    assert abs(float(sched[2]) - gr.weight_sum) <= 1e-6 * gr.weight_sum
    zr = gr.state[0]["z"]
    assert (z.cpu().double() - zr).abs().max() <= 2e-3 * zr.abs().max()
    # y is stored in bf16 on both sides: allow one rounding step of disagreement
    assert (dev_y.cpu().double() - ry[0]).abs().max() <= 2 ** -7 * ry[0].abs().max()
    # eval() / train(): the averaged iterate and back (bf16 storage: within rounding)
    before = dev_y.clone()
    ops.radam_schedulefree_swap(dev_y, z, 0.9, True)
    x_ref = ry[0] + (1 - 1 / 0.9) * (zr - ry[0])
    assert (dev_y.cpu().double() - x_ref).abs().max() <= 2 ** -6 * x_ref.abs().max()
    ops.radam_schedulefree_swap(dev_y, z, 0.9, False)
    assert (dev_y.float() - before.float()).abs().max() <= 2 ** -6 * before.float().abs().max()


@pytest.mark.parametrize("use_graph", [True, False])
def test_trainer_with_radam_schedulefree(tmp_path, use_graph):
    from vision_pt_b200 import train as T
    net = T.build_jit_qlora(_small_cfg(), rank=16, alpha=16.0, device="cuda", seed=11, lora_up_std=0.02)
    hp = T.TrainHParams(lr=2e-3, clip_grad_norm=1.0, optimizer="radam_schedulefree", weight_decay=0.0)
    tr = T.JiTQLoRATrainer(net, num_classes=10, max_token_length=16, hp=hp, use_graph=use_graph, seed=5)
    a = T.synthetic_batch(8, 64, 64, num_classes=10, max_token_length=16, seed=1)
    b = T.synthetic_batch(8, 64, 128, num_classes=10, max_token_length=16, seed=2)
    p0 = tr.state.flat.param.clone()
    lrs, losses = [], []
    for it in range(12):
        losses.append(float(tr.train_step(*(a if it % 2 == 0 else b))))
        lrs.append(tr.scheduled_lr)
        if it < 4:
            assert torch.equal(tr.state.flat.param, p0)       # RAdam's silent phase: scheduled lr = 0
    assert tr.global_step == 12 and lrs[3] == 0.0 and lrs[4] > 0.0 and lrs[-1] > lrs[4]
    assert all(torch.isfinite(torch.tensor(losses))) and not torch.equal(tr.state.flat.param, p0)
    tr.save_checkpoint(str(tmp_path))
    for batch in (a, b):
        tr.train_step(*batch)
    torch.cuda.synchronize()
    want = (tr.state.flat.param.clone(), tr.state.z.clone(), tr.state.sched.clone())
    net2 = T.build_jit_qlora(_small_cfg(), rank=16, alpha=16.0, device="cuda", seed=11, lora_up_std=0.02)
    tr2 = T.JiTQLoRATrainer(net2, num_classes=10, max_token_length=16, hp=hp, use_graph=use_graph, seed=77)
    tr2.load_checkpoint(str(tmp_path))
    for batch in (a, b):
        tr2.train_step(*batch)
    torch.cuda.synchronize()
    assert torch.equal(tr2.state.flat.param, want[0]) and torch.equal(tr2.state.z, want[1]) and torch.equal(tr2.state.sched, want[2])
    # eval(): parameters become the averaged iterate; training refuses to run until train()
    y = tr2.state.flat.param.clone()
    tr2.eval()
    assert not torch.equal(tr2.state.flat.param, y)
    with pytest.raises(RuntimeError):
        tr2.train_step(*a)
    tr2.train()
    assert (tr2.state.flat.param.float() - y.float()).abs().max() <= 2 ** -6 * y.float().abs().max()


def test_prefetched_input_pipeline_is_bit_identical():
    """train_step(batch, prefetch=next) (H2D of the next batch on a copy stream into a staging slot, device-to-device copy
    into the graph's inputs) trains exactly like the plain call (H2D on the main stream): same graph, same parameters."""
    _, plain = _trainer(True, seed=5)
    T, pref = _trainer(True, seed=5)
    shapes = [(8, 64, 64), (8, 64, 128)]
    batches = [T.synthetic_batch(B, H, W, num_classes=10, max_token_length=16, seed=i) for i, (B, H, W) in enumerate(shapes)]
    order = [0, 1, 1, 0, 0, 1]
    for tr in (plain, pref):
        tr.precapture(shapes)
    torch.manual_seed(77)
    for i in order:
        plain.train_step(*batches[i])
    torch.cuda.synchronize()
    torch.manual_seed(77)
    pref.prefetch(*batches[order[0]])
    losses = []
    for n, i in enumerate(order):
        nxt = batches[order[n + 1]] if n + 1 < len(order) else None
        pref.train_step(*batches[i], prefetch=nxt)
        if n > 0:
            losses.append(pref.read_loss(1))
    losses.append(pref.read_loss(0))
    torch.cuda.synchronize()
    assert torch.equal(plain.state.flat.param, pref.state.flat.param)
    assert torch.equal(plain.state.exp_avg, pref.state.exp_avg) and plain.global_step == pref.global_step == len(order)
    assert len(losses) == len(order) and all(l == l and l > 0 for l in losses)
