"""Training-step parity: 200 optimisation steps of the CUDA path (CUDA-graph replay, flat LoRA buffers, fused clip + AdamW)
against the oracle trained with the same batches, noise and timesteps -- BASELINE.json: "loss curve within 1 % over 200
steps".  The oracle side is fp32 autograd over oracle.jit.jit_forward with the NF4 weights dequantised by oracle.nf4 and
torch.optim-style AdamW written out below; it runs on the GPU only to finish in seconds (plain PyTorch, no kernels of
this repo)."""
import math

import pytest
import torch

from oracle import jit as oj
from oracle import nf4 as on

pytestmark = pytest.mark.gpu


def _oracle_params(net):
    P = {}
    for name, p in net.state_dict().items():
        if ".weight." in name:
            continue
        P[name] = p.detach().float()
    for name, mod in net.named_modules():
        qs = getattr(mod, "quant_state", None)
        if qs is not None:
            st = on.Nf4State(packed=qs.packed.cpu(), absmax=qs.absmax.cpu(), nested_absmax=qs.nested_absmax.cpu(),
                             nested_code=qs.nested_code.cpu(), code=qs.code.cpu(), offset=float(qs.offset),
                             shape=tuple(qs.shape), dtype=qs.dtype)
            P[f"{name}.weight"] = on.dequantize_nf4(st).float().cuda()
    return P


# "mini": a miniature net (small-M prologue GEMM route).  "jit_b_width": JiT-B/16's own widths (D 768, 12 heads, F 2048,
# bottleneck 128, context 768) at depth 4 on 256-px images, batch 8 -> M = 2640 rows: the CTA-pair GEMM / batched
# dequantisation route of real training, context tokens joining at block 1.
@pytest.mark.parametrize("which", ["mini", "jit_b_width"])
def test_loss_curve_200_steps_matches_oracle(which):
    from vision_pt_b200 import train as T
    from vision_pt_b200.jit import DenoiserConfig
    dev = torch.device("cuda")
    if which == "mini":
        cfg = DenoiserConfig(patch_size=16, in_channels=3, out_channels=3, hidden_size=128, depth=2, num_heads=2, mlp_ratio=4.0,
                             bottleneck_dim=32, num_time_tokens=4, rope_axes_dims=[16, 24, 24], context_dim=64,
                             context_start_block=1)
        B, H, W, steps = 8, 64, 64, 200
    else:
        cfg = DenoiserConfig(patch_size=16, in_channels=3, out_channels=3, hidden_size=768, depth=4, num_heads=12, mlp_ratio=4.0,
                             bottleneck_dim=128, num_time_tokens=4, rope_axes_dims=[16, 24, 24], context_dim=768,
                             context_start_block=1)
        B, H, W, steps = 8, 256, 256, 200
    cfgd = cfg.model_dump()
    hp = T.TrainHParams(lr=2e-3, clip_grad_norm=1.0, loss_target="image")
    net = T.build_jit_qlora(cfg, rank=16, alpha=16.0, device=dev, seed=11, lora_up_std=0.02)
    P = _oracle_params(net)                                   # before training: both sides start from the same weights
    step = T.JiTQLoRATrainStep(net, B, H, W, num_classes=10, max_token_length=16, hp=hp, use_graph=True, seed=5)
    host = T.synthetic_batch(B, H, W, num_classes=10, max_token_length=16, seed=3, pin=False)
    step.image.copy_(host[0]); step.class_ids.copy_(host[1]); step.attention_mask.copy_(host[2])

    # ---- CUDA path.  capture() runs 2 eager warm-up steps + the capture pass (which does not execute): replay from a
    # known state instead -- snapshot, capture, restore, then run the 200 steps.
    snap = (step.flat.param.clone(), step.exp_avg.clone(), step.exp_avg_sq.clone(), step.step_t.clone())
    step.capture()
    step.flat.param.copy_(snap[0]); step.exp_avg.copy_(snap[1]); step.exp_avg_sq.copy_(snap[2]); step.step_t.copy_(snap[3])
    step.flat.grad.zero_()
    torch.manual_seed(1234)
    ours = []
    for _ in range(steps):
        ours.append(float(step.run()))

    # ---- oracle with the same random draws (same generator state, same call sequence as JiTQLoRATrainStep._compute)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_down" in k or "lora_up" in k}
    P.update(leaves)
    m = {k: torch.zeros_like(v) for k, v in leaves.items()}
    v2 = {k: torch.zeros_like(v) for k, v in leaves.items()}
    images = host[0].to(dev)
    context = step.class_encoder(step.class_ids).float()
    mask = step.attention_mask
    size = step.size_info
    torch.manual_seed(1234)
    ref = []
    for t_step in range(1, steps + 1):
        t = (torch.randn(B, device=dev) * hp.ts_std + hp.ts_mean).sigmoid()
        noise = torch.randn_like(images) * hp.noise_scale
        tv = t.view(B, 1, 1, 1).to(images.dtype)
        noisy = tv * images + (1 - tv) * noise
        y = oj.jit_forward(P, cfgd, noisy.to(torch.bfloat16).float(), t.to(torch.bfloat16).float(), context, size, size,
                           torch.zeros_like(size), context_mask=mask, alpha=16.0)
        loss = torch.nn.functional.mse_loss(y, images.float())
        grads = torch.autograd.grad(loss, list(leaves.values()))
        ref.append(float(loss))
        total = math.sqrt(sum(float(g.pow(2).sum()) for g in grads))
        clip = min(1.0, hp.clip_grad_norm / (total + 1e-6))
        with torch.no_grad():
            for (k, p), g in zip(leaves.items(), grads):
                g = g * clip
                m[k].mul_(hp.betas[0]).add_(g, alpha=1 - hp.betas[0])
                v2[k].mul_(hp.betas[1]).addcmul_(g, g, value=1 - hp.betas[1])
                bc1, bc2 = 1 - hp.betas[0] ** t_step, 1 - hp.betas[1] ** t_step
                p.mul_(1 - hp.lr * hp.weight_decay)
                p.addcdiv_(m[k] / bc1, (v2[k] / bc2).sqrt() + hp.eps, value=-hp.lr)
                p.copy_(p.to(torch.bfloat16).float())          # the LoRA matrices are stored in bf16 (PeftConfigMixin.dtype)

    ours_t, ref_t = torch.tensor(ours), torch.tensor(ref)
    assert torch.isfinite(ours_t).all()
    # random-init frozen base + random images: the LoRA matrices can only shave a little off the loss, but they must do so
    drop_ref, drop_ours = float(ref_t[:10].mean() - ref_t[-10:].mean()), float(ours_t[:10].mean() - ours_t[-10:].mean())
    assert drop_ref > 1e-3, f"the test must actually train (loss has to fall): {ref[:3]} .. {ref[-3:]}"
    assert abs(drop_ours - drop_ref) <= 0.1 * drop_ref, (drop_ours, drop_ref)     # the same amount of learning happened
    rel = ((ours_t - ref_t).abs() / ref_t).max()
    assert rel <= 1e-2, f"loss curves differ by {float(rel):.4f} (max over 200 steps): ref {ref[:3]} .. {ref[-3:]} | ours {ours[:3]} .. {ours[-3:]}"
