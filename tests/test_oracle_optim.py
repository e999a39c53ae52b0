"""oracle/optim.py: AdamW pinned against torch.optim.AdamW; the RAdamScheduleFree restatement checked for the properties
its published algorithm has (the package itself is not in this image: parity unpinned, see the oracle's header)."""
import math

import torch

from oracle import optim as oo


def test_adamw_restatement_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(257, dtype=torch.float64)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 21):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        oo.adamw_step(p, g, m, v, step, 3e-3, (0.9, 0.95), 1e-8, 0.05)
    assert torch.allclose(p, ref.detach(), rtol=1e-12, atol=1e-14)


def test_radam_schedulefree_properties():
    torch.manual_seed(1)
    gr = oo.RAdamSFGroup(lr=1e-2)
    y = [torch.randn(64, dtype=torch.float64)]
    y0 = y[0].clone()
    lrs = []
    for _ in range(60):
        oo.radam_schedulefree_step(gr, y, [torch.randn(64, dtype=torch.float64)])
        lrs.append(gr.scheduled_lr)
        if gr.scheduled_lr == 0.0:
            assert torch.equal(y[0], y0)        # silent SGD phase (rho_t <= 4): lr = 0, nothing moves
    # RAdam rectification: zero for the first steps (rho_t <= 4 up to step 4 at beta2 = 0.999), then rising towards lr
    assert lrs[0] == 0.0 and lrs[3] == 0.0 and lrs[4] > 0.0
    assert all(b >= a for a, b in zip(lrs, lrs[1:])) and lrs[-1] < 1e-2
    rho_inf = 2 / (1 - 0.999) - 1
    rho6 = rho_inf - 2 * 6 * 0.999 ** 6 / (1 - 0.999 ** 6)
    assert math.isclose(lrs[5], 1e-2 * math.sqrt((rho6 - 4) * (rho6 - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho6)), rel_tol=1e-12)
    # eval() / train() are inverse interpolations between y and z: x = y + (1 - 1/b1)(z - y); y = x + (1 - b1)(z - x)
    before = y[0].clone()
    oo.radam_schedulefree_swap(gr, y, to_eval=True)
    assert not torch.allclose(y[0], before)
    oo.radam_schedulefree_swap(gr, y, to_eval=False)
    assert torch.allclose(y[0], before, rtol=1e-12, atol=1e-14)
    # the averaged iterate: with ckp1 = w_t / sum(w), x_t is the weighted mean of the z iterates -- check one step by hand
    gr2 = oo.RAdamSFGroup(lr=1e-2, silent_sgd_phase=False)
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64)]
    g = torch.tensor([0.5, 0.25], dtype=torch.float64)
    oo.radam_schedulefree_step(gr2, p, [g.clone()])
    # step 1: rho_t <= 4 -> plain SGD direction, rect = 1, ckp1 = 1: y = z0 + lr (b1 (1 - 1) - 1) g = z0 - lr g; z = z0 - lr g
    want = torch.tensor([1.0, -2.0], dtype=torch.float64) - 1e-2 * g
    assert torch.allclose(p[0], want) and torch.allclose(gr2.state[0]["z"], want)
