"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimisers on the hot path's update step.  Only tests/, smoke() and
bench.py's cpu_baseline may import this.

`radam_schedulefree_step` restates schedulefree.RAdamScheduleFree.step (non-foreach branch) of schedulefree 1.4.1 -- the
optimiser the reference's shipped YAMLs name (configs/jit/x-loss/config.yml:75, configs/sdxl/*.yml; pinned in
uv.lock:3428-3429; logged through param_group["scheduled_lr"], src/trainer/common.py:499-506).  The package is a
third-party dependency that is NOT in this image and not vendored in /root/reference, so this follows its published
algorithm (Defazio et al., "The Road Less Scheduled", 2024, with the RAdam rectification replacing warm-up) and is
**parity unpinned**: there is no golden vector of the package to check it against here.

`adamw_step` is torch.optim.AdamW's single-tensor update (decoupled weight decay), pinned against torch.optim.AdamW in
tests/test_oracle_golden.py."""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch


def adamw_step(p, g, m, v, step: int, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.mul_(1 - lr * weight_decay)
    p.addcdiv_(m / (1 - b1 ** step), (v / (1 - b2 ** step)).sqrt() + eps, value=-lr)


@dataclass
class RAdamSFGroup:
    """param_group state of RAdamScheduleFree (defaults of the package's constructor)."""
    lr: float = 0.0025
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 0.0
    r: float = 0.0
    weight_lr_power: float = 2.0
    silent_sgd_phase: bool = True
    k: int = 0
    lr_max: float = -1.0
    weight_sum: float = 0.0
    scheduled_lr: float = 0.0
    train_mode: bool = False
    state: dict = field(default_factory=dict)


def radam_schedulefree_coefficients(gr: RAdamSFGroup) -> dict:
    """Scalars of one step (advances the group's k / lr_max / weight_sum exactly as the package's step() does)."""
    beta1, beta2 = gr.betas
    step = gr.k + 1
    beta2_t = beta2 ** step
    bias_correction2 = 1 - beta2_t
    rho_inf = 2 / (1 - beta2) - 1                       # maximum length of the approximated SMA
    rho_t = rho_inf - 2 * step * beta2_t / bias_correction2
    if rho_t > 4.0:
        rect = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
    else:
        rect = float(not gr.silent_sgd_phase)
    lr = gr.lr * rect
    gr.scheduled_lr = lr
    gr.lr_max = max(lr, gr.lr_max)
    weight = (step ** gr.r) * (gr.lr_max ** gr.weight_lr_power)
    gr.weight_sum += weight
    ckp1 = weight / gr.weight_sum if gr.weight_sum != 0 else 0.0
    gr.k = step
    return {"lr": lr, "ckp1": ckp1, "adaptive_y_lr": lr * (beta1 * (1 - ckp1) - 1), "bias_correction2": bias_correction2,
            "adam": rho_t > 4.0}


def radam_schedulefree_step(gr: RAdamSFGroup, params: list[torch.Tensor], grads: list[torch.Tensor]) -> None:
    """One optimiser step in train mode: `params` are the y sequence (updated in place); z / exp_avg_sq live in gr.state."""
    c = radam_schedulefree_coefficients(gr)
    beta2 = gr.betas[1]
    for i, (y, grad) in enumerate(zip(params, grads)):
        st = gr.state.setdefault(i, {})
        if "z" not in st:
            st["z"] = y.clone()
            st["exp_avg_sq"] = torch.zeros_like(y)
        z, v = st["z"], st["exp_avg_sq"]
        v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        if c["adam"]:
            grad_normalized = grad / (v.div(c["bias_correction2"]).sqrt_().add_(gr.eps))
        else:
            grad_normalized = grad.clone()             # SGD (or, in the silent phase, lr = 0: nothing)
        if gr.weight_decay != 0:
            grad_normalized.add_(y, alpha=gr.weight_decay)      # weight decay calculated at y
        y.lerp_(end=z, weight=c["ckp1"])
        y.add_(grad_normalized, alpha=c["adaptive_y_lr"])
        z.sub_(grad_normalized, alpha=c["lr"])


def radam_schedulefree_swap(gr: RAdamSFGroup, params: list[torch.Tensor], to_eval: bool) -> None:
    """optimizer.eval(): y -> x = averaged iterate (p.lerp_(z, 1 - 1/beta1)); optimizer.train(): back (1 - beta1)."""
    beta1 = gr.betas[0]
    w = 1 - 1 / beta1 if to_eval else 1 - beta1
    for i, p in enumerate(params):
        if i in gr.state:
            p.lerp_(end=gr.state[i]["z"], weight=w)
