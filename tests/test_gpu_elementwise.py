"""HBM-bound kernels vs the oracle (forward and backward), plus the reference golden vectors."""
import pytest
import torch

from oracle import jit as oj
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(37, 128), (300, 768), (129, 1024), (65, 1280), (33, 2048), (10, 64)])
def test_rmsnorm(rows, D):
    from vision_pt_b200 import ops
    torch.manual_seed(D)
    x = (torch.randn(rows, D) * 2).to(torch.bfloat16)
    w = (torch.randn(D) * 0.2 + 1).to(torch.bfloat16)
    dy = torch.randn(rows, D).to(torch.bfloat16)
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    y = ops.rms_norm(xg, wg, 1e-6)
    y.backward(dy.cuda())
    assert rel_err(y, oj.rms_norm_fp32(x, w)) <= 1e-2
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    oj.rms_norm_fp32(xr, wr).backward(dy.float())
    assert rel_err(xg.grad, xr.grad) <= 2e-2 and rel_err(wg.grad, wr.grad) <= 2e-2


def test_rmsnorm_reference_vector(golden):
    from vision_pt_b200.modules.norm import get_norm_layer
    g = golden["rmsnorm"]
    m = get_norm_layer("rms", 128, eps=1e-6).to(torch.bfloat16).cuda()
    m.weight.data.copy_(g["w"])
    got = m(g["x"].cuda()).cpu()
    # one bf16 ulp at most: the kernel and ATen may round the fp32 rsqrt differently
    assert (got.float() - g["y"].float()).abs().max() <= 2 ** -7 * g["y"].float().abs().max()


@pytest.mark.parametrize("B,L,H", [(2, 50, 2), (3, 330, 12), (1, 77, 16)])
def test_qknorm_rope(B, L, H, golden):
    from vision_pt_b200 import ops
    cfg = golden["rope"]["cfg"]
    torch.manual_seed(L)
    ctx = 8
    hp = 4
    # any (height, width) whose token count is >= L works: take the table and cut it to L rows
    f = oj.rope_freqs_cis(cfg, 16 * hp, 16 * ((L + hp - 1) // hp), ctx)[:L]
    cs = torch.stack([f.real, f.imag], dim=-1).contiguous().float()
    x = torch.randn(B, L, H, 64).to(torch.bfloat16)
    w = (torch.randn(64) * 0.2 + 1).to(torch.bfloat16)
    dy = torch.randn(B, L, H, 64).to(torch.bfloat16)
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    y = ops.qknorm_rope(xg, wg, cs.cuda(), 1e-6)
    y.backward(dy.cuda())
    ref = oj.apply_rope(oj.rms_norm_fp32(x.permute(0, 2, 1, 3), w), f).permute(0, 2, 1, 3)
    assert rel_err(y, ref) <= 1e-2
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    oj.apply_rope(oj.rms_norm_fp32(xr.permute(0, 2, 1, 3), wr), f).permute(0, 2, 1, 3).backward(dy.float())
    assert rel_err(xg.grad, xr.grad) <= 2e-2 and rel_err(wg.grad, wr.grad) <= 2e-2


def test_rope_table_matches_reference(golden):
    from vision_pt_b200.jit.config import DenoiserConfig
    from vision_pt_b200.jit.denoiser import rope_table
    g = golden["rope"]
    t = rope_table(DenoiserConfig(**g["cfg"]), g["height"], g["width"], g["ctx"])
    assert torch.equal(t[..., 0], g["freqs_cis"].real) and torch.equal(t[..., 1], g["freqs_cis"].imag)


@pytest.mark.parametrize("rows,F", [(100, 2048), (33, 2730), (7, 341)])
def test_swiglu(rows, F):
    from vision_pt_b200 import ops
    torch.manual_seed(F)
    g, u, da = (torch.randn(rows, F).to(torch.bfloat16) for _ in range(3))
    gg, ug = g.cuda().requires_grad_(True), u.cuda().requires_grad_(True)
    a = ops.swiglu(gg, ug)
    a.backward(da.cuda())
    assert rel_err(a, oj.swiglu_gate(g, u)) <= 1e-2
    gr, ur = g.float().requires_grad_(True), u.float().requires_grad_(True)
    oj.swiglu_gate(gr, ur).backward(da.float())
    assert rel_err(gg.grad, gr.grad) <= 2e-2 and rel_err(ug.grad, ur.grad) <= 2e-2


@pytest.mark.parametrize("B,L,D", [(2, 37, 128), (3, 100, 768), (2, 65, 1280), (2, 130, 2560), (2, 77, 4096)])   # 2560 / 4096: CogView4 widths
def test_adaln_modulate_and_gate(B, L, D):
    from vision_pt_b200 import ops
    torch.manual_seed(D + L)
    x, h, dy = (torch.randn(B, L, D).to(torch.bfloat16) for _ in range(3))
    sc, sh, gt = (torch.randn(B, D).to(torch.bfloat16) * 0.5 for _ in range(3))
    leaves = [t.cuda().requires_grad_(True) for t in (x, sc, sh)]
    y = ops.ln_modulate(*leaves, 1e-5)
    y.backward(dy.cuda())
    assert rel_err(y, oj.adaln_modulate(x, sc, sh)) <= 1e-2
    refs = [t.float().requires_grad_(True) for t in (x, sc, sh)]
    oj.adaln_modulate(*refs).backward(dy.float())
    for a, b in zip(leaves, refs):
        assert rel_err(a.grad, b.grad) <= 2e-2
    leaves = [t.cuda().requires_grad_(True) for t in (x, h, gt)]
    y = ops.gate_residual(*leaves)
    y.backward(dy.cuda())
    assert rel_err(y, oj.gate_residual(x, h, gt)) <= 1e-2
    refs = [t.float().requires_grad_(True) for t in (x, h, gt)]
    oj.gate_residual(*refs).backward(dy.float())
    for a, b in zip(leaves, refs):
        assert rel_err(a.grad, b.grad) <= 2e-2


def test_adaln_reference_vector(golden):
    from vision_pt_b200.modules.norm import adaln_gate_residual, adaln_modulate
    g = golden["adaln"]
    y = adaln_modulate(g["x"].cuda(), g["scale"].cuda(), g["shift"].cuda(), eps=1e-6)
    assert rel_err(y, g["y"]) <= 1e-2
    assert rel_err(adaln_gate_residual(g["x"].cuda(), g["y"].cuda(), g["gate"].cuda()), g["gated"]) <= 1e-2


@pytest.mark.parametrize("B,C,H,W,p", [(1, 3, 256, 256, 4), (4, 16, 832, 1152, 2), (2, 3, 256, 256, 16), (2, 3, 512, 448, 16)])
@pytest.mark.parametrize("order", [0, 1])
def test_patchify_bit_exact(B, C, H, W, p, order):
    """reference tests/test_patch.py shapes: permutation kernels are bit-exact and inverse to each other."""
    from vision_pt_b200 import ops
    torch.manual_seed(0)
    img = torch.randn(B, C, H, W).to(torch.bfloat16)
    pt = ops.patchify_op(img.cuda(), p, order)
    assert torch.equal(pt.cpu(), oj.patchify(img, p, order))
    back = ops.unpatchify_op(pt, C, H, W, p, order)
    assert torch.equal(back.cpu(), img)


def test_patch_module_api(golden):
    from vision_pt_b200.modules.patch import patchify, unpatchify
    g = golden["patchify"]
    out = patchify(g["image"].cuda(), 16)
    assert out.latent_height == 2 and out.latent_width == 3 and torch.equal(out.patches.cpu(), g["patches"])
    assert torch.equal(unpatchify(out.patches, 2, 3, 16, 3).image.cpu(), g["image"])
    with pytest.raises(ValueError):
        patchify(torch.zeros(4, device="cuda"), 2)


@pytest.mark.parametrize("B,L,H", [(2, 70, 4), (1, 33, 3)])
def test_qknorm_rope_head_dim_80(B, L, H):
    """JiT-H (BASELINE.json configs[3]): head_dim 80, rope_axes_dims [16, 32, 32] (src/models/jit/config.py)."""
    from vision_pt_b200 import ops
    cfg = dict(patch_size=16, rope_theta=256.0, rope_axes_dims=[16, 32, 32], num_time_tokens=4)
    torch.manual_seed(L + H)
    f = oj.rope_freqs_cis(cfg, 16 * 4, 16 * ((L + 3) // 4), 8)[:L]
    cs = torch.stack([f.real, f.imag], dim=-1).contiguous().float()
    x = torch.randn(B, L, H, 80).to(torch.bfloat16)
    w = (torch.randn(80) * 0.2 + 1).to(torch.bfloat16)
    dy = torch.randn(B, L, H, 80).to(torch.bfloat16)
    xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    y = ops.qknorm_rope(xg, wg, cs.cuda(), 1e-6)
    y.backward(dy.cuda())
    ref = oj.apply_rope(oj.rms_norm_fp32(x.permute(0, 2, 1, 3), w), f).permute(0, 2, 1, 3)
    assert rel_err(y, ref) <= 1e-2
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    oj.apply_rope(oj.rms_norm_fp32(xr.permute(0, 2, 1, 3), wr), f).permute(0, 2, 1, 3).backward(dy.float())
    assert rel_err(xg.grad, xr.grad) <= 2e-2 and rel_err(wg.grad, wr.grad) <= 2e-2


@pytest.mark.parametrize("B,L,D,start", [(64, 330, 768, 266), (3, 50, 128, 10), (2, 9, 24, 4), (2, 7, 12, 3)])
def test_copy_token_slots(B, L, D, start):
    """The per-block context-slot refresh (reference denoiser.py:1092-1113) and its gradient zeroing: bit-exact against
    the torch slice assignment, including shapes that take the unaligned fallback."""
    from vision_pt_b200 import ops
    torch.manual_seed(0)
    buf = torch.randn(B, L, D, device="cuda").to(torch.bfloat16)
    tail = torch.randn(B, L - start, D, device="cuda").to(torch.bfloat16)
    want = buf.clone()
    want[:, start:] = tail
    got = buf.clone()
    ops.copy_token_slots(got, start, tail)
    assert torch.equal(got, want)
    want[:, start:] = 0
    ops.copy_token_slots(got, start, None)
    assert torch.equal(got, want)
    assert torch.equal(got[:, :start], buf[:, :start])


@pytest.mark.parametrize("rows,D,affine", [(300, 640, True), (131, 1280, True), (70, 64, False), (90, 4096, True), (64, 2560, False)])
def test_layernorm_affine(rows, D, affine):
    """nn.LayerNorm / FP32LayerNorm (src/modules/norm.py:9-17, src/models/sdxl/denoiser.py:248-250) with and without affine
    parameters, forward, dx and dw / db, narrow (register-cached) and wide (D > 2048: two-pass) rows."""
    from vision_pt_b200 import ops
    torch.manual_seed(rows + D)
    x, dy = (torch.randn(rows, D) * 2 + 0.5).to(torch.bfloat16), torch.randn(rows, D).to(torch.bfloat16)
    w = (torch.randn(D) * 0.3 + 1).to(torch.bfloat16) if affine else None
    b = (torch.randn(D) * 0.3).to(torch.bfloat16) if affine else None
    xg = x.cuda().requires_grad_(True)
    wg = w.cuda().requires_grad_(True) if affine else None
    bg = b.cuda().requires_grad_(True) if affine else None
    y = ops.layer_norm(xg, wg, bg, 1e-5)
    y.backward(dy.cuda())
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True) if affine else None
    br = b.float().requires_grad_(True) if affine else None
    yr = torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-5)
    yr.backward(dy.float())
    assert rel_err(y, yr) <= 1e-2 and rel_err(xg.grad, xr.grad) <= 2e-2
    if affine:
        assert rel_err(wg.grad, wr.grad) <= 2e-2 and rel_err(bg.grad, br.grad) <= 2e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,noise_scale,clean_at_zero", [((5, 3, 64, 64), 1.0, False), ((3, 3, 32, 48), 0.7, False),
                                                             ((4, 1, 7, 9), 1.3, True)])
def test_noise_mix_is_bit_identical_to_the_reference_ops(dtype, shape, noise_scale, clean_at_zero):
    """One kernel == prepare_scaled_noised_latents (reference src/modules/loss/flow_match.py:60-74), bit for bit: the
    reference's own function when its modules are staged (oracle/_ref), and always the op-by-op restatement."""
    from oracle import refimport
    from vision_pt_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(dtype).to(dev)
    t = torch.rand(shape[0], generator=g).to(dev)                       # fp32, as the trainer's sigmoid(randn) is
    tv = t.view(-1, 1, 1, 1).to(dtype)
    torch.manual_seed(77)
    z = torch.randn_like(x)
    noise = z * noise_scale
    want = (1 - tv) * x + tv * noise if clean_at_zero else tv * x + (1 - tv) * noise
    got, got_bf16 = ops.noise_mix(x, z, t, noise_scale, clean_at_zero=clean_at_zero)
    assert got.dtype == dtype and torch.equal(got, want)
    assert got_bf16.dtype == torch.bfloat16 and torch.equal(got_bf16, want.to(torch.bfloat16))
    if refimport.available():
        refimport.load()
        from src.modules.loss.flow_match import prepare_scaled_noised_latents
        torch.manual_seed(77)                                           # the reference draws its noise itself
        ref = prepare_scaled_noised_latents(x, tv, noise_scale=noise_scale, clean_at_zero=clean_at_zero)
        assert torch.equal(got, ref.noisy_latents)


def test_flow_loss_backward_applies_the_upstream_gradient():
    """d(3 * loss)/d pred == 3 * d loss/d pred with autograd's bf16 rounding (`dpred * dloss.to(bf16)`), both loss targets."""
    from vision_pt_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(3)
    pred0 = torch.randn(4, 3, 32, 40, device=dev).to(torch.bfloat16)
    clean = torch.randn(4, 3, 32, 40, device=dev).to(torch.float16)
    noisy = torch.randn(4, 3, 32, 40, device=dev).to(torch.float16)
    t = torch.rand(4, device=dev)
    for target in ("image", "velocity"):
        grads = []
        for k in (1.0, 3.0, 0.3):
            pred = pred0.clone().requires_grad_(True)
            (ops.flow_loss(pred, clean, noisy, t, loss_target=target) * k).backward()
            grads.append(pred.grad)
        for k, gk in zip((3.0, 0.3), grads[1:]):
            want = grads[0] * torch.tensor(k, device=dev).to(torch.bfloat16)
            assert torch.equal(gk, want)


@pytest.mark.parametrize("B,L,n,D", [(3, 50, 37, 128), (64, 330, 256, 768), (2, 9, 9, 64), (2, 11, 5, 12)])
def test_token_prefix_equals_the_slice(B, L, n, D):
    """ops.token_prefix(x, n) == x[:, :n] forward and backward (the patch tokens JiT.forward hands to the final layer,
    reference denoiser.py:1115-1124); ops.packed_tokens == .contiguous() for a slice of a wider token buffer."""
    from vision_pt_b200 import ops
    torch.manual_seed(L)
    x = torch.randn(B, L, D).to(torch.bfloat16).cuda().requires_grad_(True)
    w = torch.randn(B, n, D).to(torch.bfloat16).cuda()
    y = ops.token_prefix(x, n)
    assert y.is_contiguous() and torch.equal(y, x[:, :n])
    (y * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    (xr[:, :n] * w).sum().backward()
    assert torch.equal(x.grad, xr.grad)
    for lo, hi in ((0, n), (L - n, L)):
        view = x.detach()[:, lo:hi]
        assert torch.equal(ops.packed_tokens(view), view.contiguous()) and ops.packed_tokens(view).is_contiguous()
