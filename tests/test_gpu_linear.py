"""Fused NF4 + LoRA linear (forward, dX, LoRA gradients) through the module API vs the CPU oracle."""
import pytest
import torch

from oracle import jit as oj
from oracle import nf4 as on
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-2   # BASELINE.json: block outputs and gradients max rel err <= 2e-2 in bf16


def _make(M, K, N, rank, nf4, bias, seed=0):
    import torch.nn as nn
    from vision_pt_b200.modules.peft import LoRAConfig, LoRALinear
    from vision_pt_b200.modules.quant import NF4Linear
    torch.manual_seed(seed)
    base = nn.Linear(K, N, bias=bias).to(torch.bfloat16)
    with torch.no_grad():
        base.weight.normal_(std=0.05)
        if bias:
            base.bias.normal_(std=0.5)
    w_ref = base.weight.detach().clone()
    if nf4:
        q = NF4Linear(K, N, bias=bias)
        q.load_state_dict(base.state_dict(), assign=True)
        base = q
        w_ref = on.quantize_nf4(w_ref)
    layer = LoRALinear(LoRAConfig(rank=rank, alpha=2.0), base).cuda()
    with torch.no_grad():
        layer.lora_up.weight.normal_(std=0.05)
    x = torch.randn(M, K).to(torch.bfloat16)
    return layer, w_ref, x


@pytest.mark.parametrize("M,K,N", [(300, 128, 192), (257, 768, 768), (130, 768, 2048), (129, 2048, 768), (64, 64, 200),
                                   (200, 341, 128), (150, 128, 341), (140, 2730, 1024), (131, 1024, 2730)])
@pytest.mark.parametrize("nf4", ["prologue", "scratch", False])
@pytest.mark.parametrize("rank", [16, 4])
def test_lora_linear_forward_backward(M, K, N, nf4, rank, monkeypatch):
    """nf4 = how the NF4 weight reaches the tensor cores: per-stage prologue dequant, or once per call into the bf16
    workspace (include/vptb200.h: w_scratch); False = plain bf16 base weight."""
    from vision_pt_b200 import ops
    if nf4:
        monkeypatch.setattr(ops, "NF4_GEMM_MODE", nf4)
    layer, w_ref, x = _make(M, K, N, rank, bool(nf4), bias=True)
    if nf4:
        # the quantised module holds exactly the oracle's codes for this weight
        assert torch.equal(layer.linear.weight.cpu(), w_ref.packed)
    elif K % 8 != 0:
        pytest.skip("bf16 base weights with a ragged in_features take the unfused composition (not this kernel)")
    xg = x.cuda().requires_grad_(True)
    y = layer(xg)
    dy = torch.randn(M, N).to(torch.bfloat16)
    y.backward(dy.cuda())

    xr = x.clone().float().requires_grad_(True)
    down = layer.lora_down.weight.detach().cpu().float().requires_grad_(True)
    up = layer.lora_up.weight.detach().cpu().float().requires_grad_(True)
    wd = oj.dense_weight(w_ref).float()
    bias = layer.linear.bias.detach().cpu().float()
    yr = oj.lora_linear(xr, wd, bias, down, up, alpha=2.0)
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL
    assert rel_err(xg.grad, xr.grad) <= TOL
    assert rel_err(layer.lora_down.weight.grad, down.grad) <= TOL
    assert rel_err(layer.lora_up.weight.grad, up.grad) <= TOL
    # and against the reference's bf16 rounding order
    y_bf = oj.lora_linear(x, oj.dense_weight(w_ref), layer.linear.bias.detach().cpu(), layer.lora_down.weight.detach().cpu(),
                          layer.lora_up.weight.detach().cpu(), alpha=2.0)
    assert rel_err(y, y_bf) <= TOL


def test_zero_init_lora_equals_base():
    """reference tests/test_peft.py:99-102: with lora_up = 0 the wrapped layer returns the base output bit-for-bit."""
    import torch.nn as nn
    from vision_pt_b200.modules.peft import LoRAConfig, LoRALinear
    from vision_pt_b200.modules.quant import NF4Linear
    torch.manual_seed(3)
    lin = nn.Linear(256, 320).to(torch.bfloat16)
    q = NF4Linear(256, 320)
    q.load_state_dict(lin.state_dict(), assign=True)
    q.cuda()
    x = torch.randn(77, 256, dtype=torch.bfloat16, device="cuda")
    base_out = q(x)
    wrapped = LoRALinear(LoRAConfig(rank=16), q).cuda()
    assert torch.equal(wrapped(x), base_out)
    wrapped.set_enabled(False)
    assert torch.equal(wrapped(x), base_out)


def test_base_output_matches_dequant_matmul():
    """MatMul4Bit semantics: y = x @ dequant(W)^T + b with the bit-exact dequantised weight."""
    from vision_pt_b200.modules.quant import NF4Linear
    import torch.nn as nn
    torch.manual_seed(5)
    lin = nn.Linear(768, 768).to(torch.bfloat16)
    q = NF4Linear(768, 768)
    q.load_state_dict(lin.state_dict(), assign=True)
    q.cuda()
    x = torch.randn(200, 768, dtype=torch.bfloat16, device="cuda")
    w = q.dequantize()
    ref = (x.float() @ w.float().t() + q.bias.float())
    assert rel_err(q(x), ref) <= 1e-2


def test_scratch_and_prologue_paths_agree_bitwise():
    """Both NF4 routes feed the tensor cores the same bf16 weight bits, so whole outputs agree exactly (same tile order)."""
    from vision_pt_b200 import ops
    for (K, N) in ((768, 768), (1024, 2730), (2730, 1024)):
        layer, _, _ = _make(8, K, N, 16, True, bias=True, seed=11)
        x = (torch.randn(1500, K, device="cuda") * 0.5).to(torch.bfloat16)
        st = layer.linear.quant_state
        args = (st, layer.linear.bias, layer.lora_down.weight, layer.lora_up.weight, layer.scale)
        outs = []
        for mode in ("prologue", "scratch"):
            ops.NF4_GEMM_MODE = mode
            try:
                outs.append(ops.nf4_lora_linear(x, *args))
            finally:
                ops.NF4_GEMM_MODE = "auto"
        assert torch.equal(outs[0], outs[1]), (K, N)


@pytest.mark.parametrize("mode", ["prologue", "scratch"])
def test_residual_and_linearity(mode, monkeypatch):
    """size-independent properties at the bench shape: f(x) + r == f(x; residual=r); f(2x) - b == 2 (f(x) - b)."""
    from vision_pt_b200 import ops
    monkeypatch.setattr(ops, "NF4_GEMM_MODE", mode)
    layer, _, _ = _make(8, 768, 768, 16, True, bias=True, seed=9)
    M = 21120
    x = (torch.randn(M, 768, device="cuda") * 0.5).to(torch.bfloat16)
    r = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
    st = layer.linear.quant_state
    args = (st, layer.linear.bias, layer.lora_down.weight, layer.lora_up.weight, layer.scale)
    y = ops.nf4_lora_linear(x, *args)
    yr = ops.nf4_lora_linear(x, *args, residual=r)
    assert rel_err(yr, y.float() + r.float()) <= 1e-2
    y2 = ops.nf4_lora_linear(x * 2, *args)
    b = layer.linear.bias.float()
    assert rel_err(y2.float() - b, 2 * (y.float() - b)) <= 1e-2


@pytest.mark.parametrize("K,N,bias", [(640, 640, False), (640, 640, True), (2048, 640, False), (640, 5120, True), (2560, 640, True),
                                      (1280, 1280, False), (2048, 1280, False), (1280, 10240, True), (5120, 1280, True)])
@pytest.mark.parametrize("rank", [4, 16])
def test_sdxl_transformer_block_linears(K, N, bias, rank):
    """BASELINE.json configs[4] / SURVEY 8a12: the NF4 + LoRA linears of an SDXL TransformerBlock (attn1.to_q/k/v and attn2.to_q
    without bias, attn2.to_k/v from the 2048-wide context, to_out.0 with bias, GEGLU ff.net.0.proj D -> 8D, ff.net.2 4D -> D;
    src/models/sdxl/denoiser.py:32-207) at D = 640 and 1280, LoRA ranks 4 and 16 (configs/sdxl/*.yml)."""
    M = 1024 + 77
    layer, w_ref, x = _make(M, K, N, rank, True, bias=bias, seed=K + N)
    xg = x.cuda().requires_grad_(True)
    y = layer(xg)
    dy = torch.randn(M, N).to(torch.bfloat16)
    y.backward(dy.cuda())
    xr = x.clone().float().requires_grad_(True)
    down = layer.lora_down.weight.detach().cpu().float().requires_grad_(True)
    up = layer.lora_up.weight.detach().cpu().float().requires_grad_(True)
    b = layer.linear.bias.detach().cpu().float() if bias else None
    yr = oj.lora_linear(xr, oj.dense_weight(w_ref).float(), b, down, up, alpha=2.0)
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL
    assert rel_err(xg.grad, xr.grad) <= TOL
    assert rel_err(layer.lora_down.weight.grad, down.grad) <= TOL
    assert rel_err(layer.lora_up.weight.grad, up.grad) <= TOL


@pytest.mark.parametrize("K,N", [(768, 768), (768, 2048), (2048, 768)])
def test_full_size_jit_b_linear(K, N):
    """BASELINE.json configs[1] size: M = 64 x 330 = 21120 rows (the CTA-pair route, several waves of tiles, a ragged last
    row tile) -- forward, dX and the LoRA gradients against the oracle's formula evaluated in fp32 on the same device, plus
    the size-independent properties: zero-initialised lora_up leaves the base output untouched, and rows are independent
    (any row slice of the big call equals the same rows computed alone)."""
    M = 21120
    layer, w_ref, x = _make(M, K, N, 16, True, bias=True, seed=K + N)
    xg = x.cuda().requires_grad_(True)
    y = layer(xg)
    dy = torch.randn(M, N).to(torch.bfloat16).cuda()
    y.backward(dy)
    wd = oj.dense_weight(w_ref).float().cuda()
    xr = x.cuda().float().requires_grad_(True)
    down = layer.lora_down.weight.detach().float().requires_grad_(True)
    up = layer.lora_up.weight.detach().float().requires_grad_(True)
    yr = oj.lora_linear(xr, wd, layer.linear.bias.detach().float(), down, up, alpha=2.0)
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL and rel_err(xg.grad, xr.grad) <= TOL
    assert rel_err(layer.lora_down.weight.grad, down.grad) <= TOL and rel_err(layer.lora_up.weight.grad, up.grad) <= TOL
    with torch.no_grad():
        rows = slice(20991, 21120)                         # the ragged tail of the last 256-row tile
        y_tail = layer(x[rows].cuda())
        assert rel_err(y_tail, y[rows]) <= 1e-6 + 2 ** -7   # same arithmetic, possibly another route (prologue vs pair)
        layer.lora_up.weight.zero_()
        y0 = layer(x.cuda())
        yb = layer.linear(x.cuda())
        assert torch.equal(y0, yb)


@pytest.mark.parametrize("M,K,N", [(300, 3413, 1280), (17568, 3413, 1280), (21120, 2730, 1024), (130, 341, 128),
                                   (300, 1280, 3413), (16384, 1024, 2730), (130, 128, 341), (16384, 768, 768)])
def test_frozen_bf16_linear_with_ragged_in_features(M, K, N):
    """The final layer's SwiGLU w_3 of JiT-H / JiT-L (frozen bf16, in_features 3413 / 2730: not a multiple of 8) runs the
    tcgen05 kernels through a row-padded view of the weight (ops.padded_weight) in both directions, small and large M."""
    from vision_pt_b200 import ops
    from vision_pt_b200.jit.denoiser import _kernel_linear
    torch.manual_seed(K)
    w = (torch.randn(N, K) * 0.03).to(torch.bfloat16).cuda()
    b = (torch.randn(N) * 0.2).to(torch.bfloat16).cuda()
    x = torch.randn(M, K).to(torch.bfloat16).cuda().requires_grad_(True)
    y = _kernel_linear(x, w, b)
    assert y is not None and y.shape == (M, N)
    dy = torch.randn(M, N).to(torch.bfloat16).cuda()
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    yr = xr @ w.float().t() + b.float()
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL and rel_err(x.grad, xr.grad) <= TOL
    pw = ops.padded_weight(w)
    assert pw.shape == (N, K) and pw.stride(0) % 8 == 0 and torch.equal(pw, w) and ops.padded_weight(w) is pw   # cached
    # the input gradient ran as a forward call on the cached transpose (CTA-pair kernel); the 1-CTA backward kernel agrees
    tw = ops.transposed_weight(pw)
    assert tw.shape == (K, N) and tw.stride(0) % 8 == 0 and torch.equal(tw, w.t()) and ops.transposed_weight(pw) is tw
    ops.DENSE_BWD_VIA_TRANSPOSE = False
    try:
        x2 = x.detach().clone().requires_grad_(True)
        _kernel_linear(x2, w, b).backward(dy)
    finally:
        ops.DENSE_BWD_VIA_TRANSPOSE = True
    assert rel_err(x2.grad, xr.grad) <= TOL and rel_err(x2.grad, x.grad) <= TOL


@pytest.mark.parametrize("M,D,F", [(2640, 768, 2048), (1320, 1024, 2730), (1100, 1280, 3413), (21120, 768, 2048)])
def test_swiglu_epilogues_equal_the_separate_kernels(M, D, F):
    """SwiGLU forward fused into the w_2 GEMM's epilogue and SwiGLU backward fused into the w_3 dX GEMM's epilogue (CTA-pair
    route, include/vptb200.h epilogue = 1 / 2; reference SwiGLU.forward, jit/denoiser.py:498-506) against the same GEMMs
    followed by the stand-alone swiglu kernels, and against the fp32 oracle formula.  Ragged F (JiT-L 2730, JiT-H 3413)."""
    from vision_pt_b200 import ops
    _, _, _ = _make(8, 64, 64, 16, True, True)          # library loaded
    g_ = torch.Generator().manual_seed(M + F)
    rnd = lambda *s, std=1.0: (torch.randn(*s, generator=g_) * std).to(torch.bfloat16)
    code = torch.tensor(on.NF4_CODE.copy())
    ncode = torch.from_numpy(on.dynamic_map_signed8().copy())
    w2 = ops.nf4_quantize(rnd(F, D, std=0.05).cuda(), ncode, code)
    w3 = ops.nf4_quantize(rnd(D, F, std=0.05).cuda(), ncode, code)
    b2 = rnd(F, std=0.5).cuda()
    down2, up2 = rnd(16, D, std=0.05).cuda(), rnd(F, 16, std=0.05).cuda()
    down3, up3 = rnd(16, F, std=0.05)[:, :F].cuda(), rnd(D, 16, std=0.05).cuda()
    down3 = ops._pad_rank(down3, up3)[0]
    h, g = rnd(M, D).cuda(), ops._rows(rnd(M, F).cuda())
    dy = rnd(M, D).cuda()
    slots_f = ops.dequant_block([w2, w3], [down2, down3], [up2, up3], transposed=False)
    # ---- forward: u = w_2(h), a = silu(g) * u
    u_ref, _ = ops.linear_raw(h, w2, b2, down2, up2, 0.5, want_side=True, scratch=slots_f[0])
    a_ref = ops.swiglu_fwd_raw(g, u_ref)
    a, u, side = ops.linear_raw(h, w2, b2, down2, up2, 0.5, g, want_side=True, scratch=slots_f[0], epilogue=1)
    assert torch.equal(u, u_ref) and side is not None
    gf, uf = g.float(), u_ref.float()
    a_oracle = oj.swiglu_gate(gf, uf)
    assert rel_err(a, a_oracle) <= 1e-2 and rel_err(a, a_ref) <= 8e-3      # one bf16 ulp at most (approximate divide)
    assert float((a != a_ref).float().mean()) < 2e-3
    # ---- backward: da = dy W_3, (dg, du) = swiglu'(da, g, u)
    slots_b = ops.dequant_block([w2, w3], [down2, down3], [up2, up3], transposed=True)
    da_ref, dside_ref = ops.linear_raw(dy, w3, None, down3, up3, 0.5, want_side=True, backward=True, scratch=slots_b[1])
    dg_ref, du_ref = ops.swiglu_bwd_raw(da_ref, g, u_ref)
    dg, du, dside = ops.linear_raw(dy, w3, None, down3, up3, 0.5, g, want_side=True, backward=True, scratch=slots_b[1],
                                   epilogue=2, in2=u_ref)
    assert torch.equal(dside[:, :M], dside_ref[:, :M])
    daf = da_ref.float()
    sg = torch.sigmoid(gf)
    assert rel_err(du, daf * gf * sg) <= 1e-2 and rel_err(dg, daf * uf * (sg * (1 + gf * (1 - sg)))) <= 1e-2
    assert rel_err(du, du_ref) <= 8e-3 and rel_err(dg, dg_ref) <= 8e-3
    assert float((du != du_ref).float().mean()) < 2e-3 and float((dg != dg_ref).float().mean()) < 2e-3


@pytest.mark.parametrize("rank", [32, 64])
def test_lora_rank_above_16_takes_the_composition(rank):
    """LoRAConfig.rank is free in the reference (src/modules/peft/lora.py:11-16).  The fused kernel folds ranks up to 16 into
    the GEMM; larger ranks run the reference's composition -- the NF4 base through the fused kernel, the two skinny LoRA
    GEMMs, scale and add as separate launches -- with the same result."""
    layer, w_ref, x = _make(300, 768, 768, rank, True, bias=True)
    assert not layer.fusable
    xg = x.cuda().requires_grad_(True)
    y = layer(xg)
    dy = torch.randn(300, 768).to(torch.bfloat16)
    y.backward(dy.cuda())
    xr = x.clone().float().requires_grad_(True)
    down = layer.lora_down.weight.detach().cpu().float().requires_grad_(True)
    up = layer.lora_up.weight.detach().cpu().float().requires_grad_(True)
    yr = oj.lora_linear(xr, oj.dense_weight(w_ref).float(), layer.linear.bias.detach().cpu().float(), down, up, alpha=2.0)
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL and rel_err(xg.grad, xr.grad) <= TOL
    assert rel_err(layer.lora_down.weight.grad, down.grad) <= TOL and rel_err(layer.lora_up.weight.grad, up.grad) <= TOL


@pytest.mark.parametrize("M,D,F", [(300, 128, 341), (16384, 768, 2048), (4000, 1024, 2730), (1100, 1280, 3413)])
def test_dense_swiglu_fused_equals_composed(M, D, F):
    """The final layer's MLP (three frozen bf16 linears, reference jit/denoiser.py:498-506, 535-543) with the gate in the GEMM
    epilogues (ops.dense_swiglu) against the composition linear -> swiglu kernel -> linear and against fp32 torch, forward
    and input gradient; ragged F as in JiT-L / JiT-H."""
    from vision_pt_b200 import ops
    from vision_pt_b200.jit.denoiser import SwiGLU
    torch.manual_seed(M + F)
    mlp = SwiGLU(D, int(F * 3 / 2) + 1, bias=True)
    assert mlp.w_1.out_features == F
    for p in mlp.parameters():
        p.data.normal_(0, 0.05)
    mlp = mlp.to(torch.bfloat16).cuda().requires_grad_(False)
    x = torch.randn(M, D).to(torch.bfloat16).cuda().requires_grad_(True)
    dy = torch.randn(M, D).to(torch.bfloat16).cuda()
    assert ops.FUSE_SWIGLU
    y = mlp(x)
    y.backward(dy)
    ops.FUSE_SWIGLU = False
    try:
        xc = x.detach().clone().requires_grad_(True)
        yc = mlp(xc)
        yc.backward(dy)
    finally:
        ops.FUSE_SWIGLU = True
    xr = x.detach().float().requires_grad_(True)
    w = {k: v.float() for k, v in mlp.state_dict().items()}
    g = xr @ w["w_1.weight"].t() + w["w_1.bias"]
    u = xr @ w["w_2.weight"].t() + w["w_2.bias"]
    yr = (torch.nn.functional.silu(g) * u) @ w["w_3.weight"].t() + w["w_3.bias"]
    yr.backward(dy.float())
    assert rel_err(y, yr) <= TOL and rel_err(x.grad, xr.grad) <= TOL
    assert rel_err(y, yc) <= 1e-2 and rel_err(x.grad, xc.grad) <= 1e-2
