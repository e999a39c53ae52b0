"""Data-parallel host logic on CPU, world size 2, gloo: the flat LoRA-gradient buffer, its single SUM all-reduce and the
1/world factor reproduce the gradient of the full batch (what DDP gives the reference: src/trainer/common.py:62-65,
376-380).  The per-rank gradients come from the CPU oracle; the CUDA kernels are not involved here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(32, 48)
        self.b = nn.Linear(48, 24)     # 24*16 and 16*48: slices that are multiples of 8 anyway; a.lora has 16*32


def _build(seed: int):
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig
    torch.manual_seed(seed)
    m = _Toy().to(torch.bfloat16)
    m.requires_grad_(False)
    PeftTargetConfig(include_keys=["a", "b"], config=LoRAConfig(rank=16, alpha=8.0)).replace_to_peft_layer(m)
    for n, p in m.named_parameters():
        if "lora_up" in n:
            nn.init.normal_(p, std=0.05)
        p.requires_grad_("lora_" in n)
    return m


def _oracle_grads(m, x, y):
    """mean-squared-error loss of the two LoRA linears, fp32 on the CPU oracle; returns {param: grad}."""
    from oracle import jit as oj
    leaves = {}

    def lin(layer, inp):
        d = layer.lora_down.weight.detach().float().requires_grad_(True)
        u = layer.lora_up.weight.detach().float().requires_grad_(True)
        leaves[layer.lora_down.weight] = d
        leaves[layer.lora_up.weight] = u
        return oj.lora_linear(inp, layer.linear.weight.float(), layer.linear.bias.float(), d, u, alpha=float(layer.alpha))

    out = lin(m.b, torch.tanh(lin(m.a, x)))
    loss = torch.nn.functional.mse_loss(out, y)
    loss.backward()
    return {p: t.grad for p, t in leaves.items()}, float(loss)


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vision_pt_b200.train import FlatLoRA
        m = _build(seed=3)                                   # replicas: same seed on every rank
        flat = FlatLoRA(m)
        g = torch.Generator().manual_seed(11)
        X, Y = torch.randn(8, 32, generator=g), torch.randn(8, 24, generator=g)
        per = 8 // world
        xs, ys = X[rank * per:(rank + 1) * per], Y[rank * per:(rank + 1) * per]   # the batch is what is sharded
        grads, _ = _oracle_grads(m, xs, ys)
        for p, gr in grads.items():
            p._vpt_grad32.add_(gr)                           # what the lora_grad kernels do on the GPU
        scale = flat.all_reduce(dist.group.WORLD)
        q.put((rank, scale, flat.grad.clone(), [int(o) for o in flat.offsets], flat.numel))
    finally:
        dist.destroy_process_group()


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_flat_lora_allreduce_world2_matches_full_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, g0, off0, n0), (_, s1, g1, off1, n1) = res
    assert s0 == s1 == 0.5 and off0 == off1 and n0 == n1
    assert torch.equal(g0, g1), "ranks disagree after the all-reduce"
    assert all(o % 8 == 0 for o in off0), "every slice of the flat buffer must start on a 16-byte boundary"

    # single-process truth: gradient of the mean loss over the full batch = mean of the per-rank gradients
    from vision_pt_b200.train import FlatLoRA
    m = _build(seed=3)
    flat = FlatLoRA(m)
    g = torch.Generator().manual_seed(11)
    X, Y = torch.randn(8, 32, generator=g), torch.randn(8, 24, generator=g)
    grads, _ = _oracle_grads(m, X, Y)
    for p, gr in grads.items():
        p._vpt_grad32.add_(gr)
    assert flat.all_reduce(None) == 1.0                      # no process group: identity
    torch.testing.assert_close(g0 * s0, flat.grad, rtol=1e-5, atol=1e-7)


def test_flat_lora_views_alias_the_parameters():
    from vision_pt_b200.train import FlatLoRA
    m = _build(seed=5)
    before = {n: p.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
    flat = FlatLoRA(m)
    assert flat.numel == sum((p.numel() + 7) // 8 * 8 for p in flat.params)
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert torch.equal(p, before[n])                 # values survived the move into the flat buffer
            assert p.data_ptr() >= flat.param.data_ptr() and p.data_ptr() < flat.param.data_ptr() + flat.numel * 2
            assert p._vpt_grad32.shape == p.shape and p._vpt_grad32.dtype == torch.float32
    flat.param.zero_()
    assert all(float(p.abs().sum()) == 0 for p in flat.params)   # one buffer: an optimiser kernel over it updates all
