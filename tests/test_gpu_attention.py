"""sm_100a attention forward / backward vs the oracle (explicit fp32 softmax) and the reference golden vector."""
import pytest
import torch

from oracle import jit as oj
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,Lq,Lk,lens", [(2, 3, 330, 330, [330, 285]), (1, 2, 128, 128, None), (2, 2, 266, 266, None),
                                           (2, 1, 200, 77, [77, 40]), (1, 2, 1100, 1100, [1093]),
                                           # SDXL TransformerBlock at D=1280 (20 heads of 64): self-attention over ~1024 latent
                                           # tokens and cross-attention to 77 / 231 text tokens (src/models/sdxl/denoiser.py:32-172)
                                           (1, 20, 1056, 1056, None), (2, 20, 1024, 231, None), (1, 10, 2112, 77, None),
                                           # edges: fewer queries than one sub-tile, a single valid key, ragged tails
                                           (1, 1, 40, 24, None), (2, 2, 17, 330, [330, 1]), (3, 1, 129, 129, [129, 65, 16]),
                                           (1, 3, 64, 640, [577])])
@pytest.mark.parametrize("layout", ["bhld", "blhd"])
def test_attention_fwd_bwd(B, H, Lq, Lk, lens, layout):
    from vision_pt_b200 import ops
    torch.manual_seed(B * 100 + Lq)
    mk = lambda L: torch.randn(B, H, L, 64).to(torch.bfloat16)
    q, k, v, d_o = mk(Lq) * 1.5, mk(Lk) * 1.5, mk(Lk), mk(Lq)
    seq = None if lens is None else torch.tensor(lens, dtype=torch.int32)

    def dev(t):
        if layout == "bhld":
            return t.cuda().requires_grad_(True)
        return t.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3).requires_grad_(True)

    qg, kg, vg = dev(q), dev(k), dev(v)
    o = ops.attention(qg, kg, vg, None if seq is None else seq.cuda())
    o.backward(d_o.cuda())
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    orf = oj.attention_explicit(qr, kr, vr, None if seq is None else seq.long())
    orf.backward(d_o.float())
    assert rel_err(o, orf) <= 2e-2
    assert rel_err(qg.grad, qr.grad) <= 2e-2
    assert rel_err(kg.grad, kr.grad) <= 2e-2
    assert rel_err(vg.grad, vr.grad) <= 2e-2
    if seq is not None:   # padded keys receive exactly zero gradient
        for b, n in enumerate(lens):
            if n < Lk:
                assert kg.grad[b, :, n:].abs().max() == 0 and vg.grad[b, :, n:].abs().max() == 0


def test_reference_sdpa_vector(golden):
    """scaled_dot_product_attention(q, k, v, mask=[B,H,L,L] expanded key padding) of the reference, same call."""
    from vision_pt_b200.modules.attention import scaled_dot_product_attention
    g = golden["attention"]
    B, H, L, _ = g["q"].shape
    mask = g["key_mask"].bool().cuda().view(B, 1, 1, L).expand(-1, H, L, -1)   # what JiT's Attention builds
    y = scaled_dot_product_attention(g["q"].cuda(), g["k"].cuda(), g["v"].cuda(), mask=mask)
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) <= 2e-2
    with pytest.raises(NotImplementedError):
        scaled_dot_product_attention(g["q"].cuda(), g["k"].cuda(), g["v"].cuda(), mask=torch.ones(B, H, L, L, dtype=torch.bool, device="cuda"))
    with pytest.raises(ValueError):
        scaled_dot_product_attention(g["q"].cuda(), g["k"].cuda(), g["v"].cuda(), backend="nope")


@pytest.mark.parametrize("B,H,Lq,Lk,lens,hd", [(2, 3, 300, 300, [300, 251], 80), (1, 2, 1100, 1100, [1093], 80),
                                              (2, 2, 200, 77, [77, 40], 80), (1, 2, 130, 130, None, 128), (1, 1, 64, 40, None, 32),
                                              # head_dim 80 runs the tcgen05 kernels on two-block tiles: the same edges as head_dim 64
                                              (1, 1, 40, 24, None, 80), (2, 2, 17, 330, [330, 1], 80), (3, 1, 129, 129, [129, 65, 16], 80),
                                              (1, 3, 64, 640, [577], 80), (16, 16, 1098, 1098, None, 80)])
def test_attention_other_head_dims(B, H, Lq, Lk, lens, hd):
    """head_dim 80 (JiT-H, BASELINE.json configs[3]; the last row is its full size) runs the tcgen05 kernels templated on
    head_dim; 32 / 128 run the CUDA-core kernels of attention_simple.cuh: same contract and tolerance everywhere."""
    from vision_pt_b200 import ops
    torch.manual_seed(B * 100 + Lq + hd)
    mk = lambda L: torch.randn(B, H, L, hd).to(torch.bfloat16)
    q, k, v, d_o = mk(Lq) * 1.5, mk(Lk) * 1.5, mk(Lk), mk(Lq)
    seq = None if lens is None else torch.tensor(lens, dtype=torch.int32)
    dev = lambda t: t.permute(0, 2, 1, 3).contiguous().cuda().permute(0, 2, 1, 3).requires_grad_(True)
    qg, kg, vg = dev(q), dev(k), dev(v)
    o = ops.attention(qg, kg, vg, None if seq is None else seq.cuda())
    o.backward(d_o.cuda())
    big = B * H * Lq * Lk > 1 << 24               # the full-size row: evaluate the oracle on the GPU
    put = (lambda t: t.cuda()) if big else (lambda t: t)
    qr, kr, vr = (put(t).float().requires_grad_(True) for t in (q, k, v))
    orf = oj.attention_explicit(qr, kr, vr, None if seq is None else put(seq).long())
    orf.backward(put(d_o).float())
    assert rel_err(o, orf) <= 2e-2
    assert rel_err(qg.grad, qr.grad) <= 2e-2 and rel_err(kg.grad, kr.grad) <= 2e-2 and rel_err(vg.grad, vr.grad) <= 2e-2
    if seq is not None:
        for b, n in enumerate(lens):
            if n < Lk:
                assert kg.grad[b, :, n:].abs().max() == 0 and vg.grad[b, :, n:].abs().max() == 0


def test_attention_full_size_jit_b():
    """BASELINE.json configs[1] shape (B=64, H=12, L=330, per-sample valid key lengths as the class-label context gives
    them) against the oracle evaluated in fp32 on the same device."""
    from vision_pt_b200 import ops
    torch.manual_seed(7)
    B, H, L = 64, 12, 330
    mk = lambda: torch.randn(B, L, H, 64, device="cuda").to(torch.bfloat16).permute(0, 2, 1, 3)
    q, k, v, d_o = mk() * 1.5, mk() * 1.5, mk(), mk()
    seq = torch.randint(L - 56, L + 1, (B,), device="cuda", dtype=torch.int32)
    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = ops.attention(qg, kg, vg, seq)
    o.backward(d_o)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    orf = oj.attention_explicit(qr, kr, vr, seq.long())
    orf.backward(d_o.float())
    assert rel_err(o, orf) <= 2e-2
    assert rel_err(qg.grad, qr.grad) <= 2e-2 and rel_err(kg.grad, kr.grad) <= 2e-2 and rel_err(vg.grad, vr.grad) <= 2e-2
    for b in range(0, B, 7):
        n = int(seq[b])
        if n < L:
            assert kg.grad[b, :, n:].abs().max() == 0 and vg.grad[b, :, n:].abs().max() == 0
