"""JiT block and whole denoiser on the B200 kernels vs the oracle / the reference golden vectors."""
import pytest
import torch

from oracle import jit as oj
from oracle import nf4 as on
from tests.helpers import lora_param_dict, rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _build_model(cfgd, base_state, lora_state=None, alpha=4.0, nf4=False, dtype=torch.bfloat16):
    """Our Denoiser with the reference's weights (optionally NF4-quantised block linears and LoRA adapters)."""
    from vision_pt_b200.jit import Denoiser, DenoiserConfig
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig
    from vision_pt_b200.modules.quant import quantize_inplace
    from vision_pt_b200.modules.state_dict import RegexMatch
    model = Denoiser(DenoiserConfig(**cfgd))
    model.load_state_dict(base_state)
    model.to(dtype)
    model.requires_grad_(False)
    if nf4:
        names = [n for n, _ in model.named_modules()]
        keys = [n for n in names if n.startswith("blocks.") and n.split(".")[-1] in ("to_q", "to_k", "to_v", "to_o", "w_1", "w_2", "w_3")]
        quantize_inplace(model, "bnb_nf4", keys)
    if lora_state is not None:
        PeftTargetConfig(include_keys=[RegexMatch(regex=r"blocks\.\d+\.(attn|mlp)\.")],
                         config=LoRAConfig(rank=16, alpha=alpha)).replace_to_peft_layer(model)
        for name, mod in model.named_modules():
            if hasattr(mod, "load_weights") and f"{name}.lora_down.weight" in lora_state:
                mod.load_weights({k: lora_state.get(f"{name}.{k}") for k in mod.adapter_weight_names if k != "alpha"})
                mod.lora_down.weight.data = mod.lora_down.weight.data.to(dtype)
                mod.lora_up.weight.data = mod.lora_up.weight.data.to(dtype)
        for n, p in model.named_parameters():
            p.requires_grad_("lora_down" in n or "lora_up" in n)
    return model.cuda()


def test_block_reference_vector(golden):
    """One JiTBlock with the reference's weights.  (1) bf16 weights, kernels composed by autograd vs the reference output;
    (2) the same block NF4-quantised (hidden 341 -> ragged repack path), fused sequence vs per-op vs the oracle run on the
    dequantised weights."""
    from vision_pt_b200.jit import DenoiserConfig, JiTBlock
    from vision_pt_b200.jit.denoiser import rope_table
    from vision_pt_b200.modules.quant import quantize_inplace
    g = golden["block_bf16"]
    cfg = DenoiserConfig(**golden["rope"]["cfg"])
    blk = JiTBlock(hidden_dim=128, num_heads=2)
    blk.load_state_dict(g["state"])
    blk.to(torch.bfloat16).cuda().requires_grad_(False)
    cs = rope_table(cfg, golden["rope"]["height"], golden["rope"]["width"], golden["rope"]["ctx"]).cuda()
    seq = g["key_mask"].sum(1).to(torch.int32).cuda()
    x = g["x"].cuda()
    valid = g["key_mask"].bool()
    y_ops = blk(x, cs, seq)
    for b in range(x.shape[0]):      # padded context rows are never keys and are stripped by the model: valid rows only
        assert rel_err(y_ops[b][valid[b]], g["y"][b][valid[b]]) <= TOL

    blk.cpu()
    quantize_inplace(blk, "bnb_nf4", ["attn.to_", "mlp.w_"])
    blk.cuda()
    assert blk.fused_eligible(x)
    y_fused = blk(x, cs, seq)
    blk.use_fused = False
    y_ops = blk(x, cs, seq)
    P = dict(g["state"])
    for k in list(P):
        if k.endswith(".weight") and P[k].dim() == 2:
            P[k] = on.dequantize_nf4(on.quantize_nf4(P[k]))
    f = oj.rope_freqs_cis(golden["rope"]["cfg"], golden["rope"]["height"], golden["rope"]["width"], golden["rope"]["ctx"])
    y_ref = oj.jit_block(P, "", g["x"], f, g["key_mask"], num_heads=2)
    for b in range(x.shape[0]):
        assert rel_err(y_fused[b][valid[b]], y_ref[b][valid[b]]) <= TOL
        assert rel_err(y_ops[b][valid[b]], y_ref[b][valid[b]]) <= TOL
    assert rel_err(y_fused, y_ops) <= 1e-2


@pytest.mark.parametrize("nf4", [False, True])
def test_denoiser_lora_grads(golden, nf4):
    """Whole denoiser with LoRA (rank 16) on every block linear: prediction, loss and every LoRA gradient vs the oracle
    evaluated in fp32 on the same (bf16-rounded, optionally NF4-dequantised) weights."""
    g, base = golden["denoiser_lora_f32"], golden["denoiser_f32"]["state"]
    cfgd = g["cfg"]
    model = _build_model(cfgd, base, g["lora_state"], alpha=g["alpha"], nf4=nf4)
    inp = g["inputs"]
    kw = dict(image=inp["image"].to(torch.bfloat16).cuda(), timestep=inp["timestep"].to(torch.bfloat16).cuda(),
              context=inp["context"].to(torch.bfloat16).cuda(), original_size=inp["original_size"].cuda(),
              target_size=inp["target_size"].cuda(), crop_coords=inp["crop_coords"].cuda(), context_mask=inp["context_mask"].cuda())
    pred = model(**kw)
    tt = inp["timestep"].to(torch.bfloat16).float().cuda()
    loss = oj.velocity_loss(pred.float(), g["clean"].cuda(), inp["image"].to(torch.bfloat16).float().cuda(), tt)
    loss.backward()

    # oracle on the same effective weights, fp32 math
    rnd = lambda t: t.to(torch.bfloat16).float() if t.is_floating_point() else t
    P = {}
    for k, v in lora_param_dict(base, g["lora_state"]).items():
        v = rnd(v)
        if nf4 and k.startswith("blocks.") and k.endswith(".linear.weight"):
            v = on.dequantize_nf4(on.quantize_nf4(v.to(torch.bfloat16))).float()
        P[k] = v
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_" in k}
    P.update(leaves)
    o_inp = {k: rnd(v) for k, v in inp.items()}
    y = oj.jit_forward(P, cfgd, **o_inp, alpha=g["alpha"])
    ref_loss = oj.velocity_loss(y, g["clean"], o_inp["image"], o_inp["timestep"])
    ref_loss.backward()

    assert rel_err(pred, y) <= 3e-2                     # 2 blocks + final layer in bf16 vs fp32
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * abs(float(ref_loss))
    worst = 0.0
    for name, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, name
            worst = max(worst, rel_err(p.grad, leaves[name].grad))
    assert worst <= 5e-2, worst                          # bf16 end-to-end through the whole network
    if not nf4:
        # the fp32 reference run itself (golden) agrees with the oracle-on-rounded-weights up to the rounding of weights
        assert abs(float(ref_loss) - float(g["loss"])) <= 5e-2 * abs(float(g["loss"]))


def test_denoiser_plain_reference_vector(golden):
    g = golden["denoiser_f32"]
    model = _build_model(g["cfg"], g["state"])
    inp = g["inputs"]
    with torch.no_grad():
        y = model(image=inp["image"].to(torch.bfloat16).cuda(), timestep=inp["timestep"].to(torch.bfloat16).cuda(),
                  context=inp["context"].to(torch.bfloat16).cuda(), original_size=inp["original_size"].cuda(),
                  target_size=inp["target_size"].cuda(), crop_coords=inp["crop_coords"].cuda(),
                  context_mask=inp["context_mask"].cuda())
    assert y.shape == g["y"].shape and rel_err(y, g["y"]) <= 3e-2


def test_fused_block_equals_per_op_block_at_bench_shape():
    """JiT-B block at batch 8 x 330 tokens, NF4 + LoRA: fused sequence vs autograd-composed kernels (outputs + grads)."""
    from vision_pt_b200.jit import DenoiserConfig, JiTBlock, JiT_B_16_Config
    from vision_pt_b200.jit.denoiser import rope_table
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig
    from vision_pt_b200.modules.quant import quantize_inplace
    torch.manual_seed(0)
    cfg = JiT_B_16_Config()
    blk = JiTBlock(768, 12)
    for n, p in blk.named_parameters():
        torch.nn.init.normal_(p, std=0.02) if p.dim() > 1 else torch.nn.init.normal_(p, mean=(1.0 if "norm" in n else 0.0), std=0.05)
    blk.to(torch.bfloat16).requires_grad_(False)
    quantize_inplace(blk, "bnb_nf4", ["attn.to_", "mlp.w_"])
    PeftTargetConfig(include_keys=["attn.to_", "mlp.w_"], config=LoRAConfig(rank=16, alpha=16.0)).replace_to_peft_layer(blk)
    blk.cuda()
    for n, p in blk.named_parameters():
        if "lora_up" in n:
            torch.nn.init.normal_(p, std=0.02)
        p.requires_grad_("lora_" in n and "alpha" not in n)
    B, L = 8, 330
    cs = rope_table(cfg, 256, 256, 64).cuda()
    seq = torch.randint(266 + 8, 331, (B,), dtype=torch.int32).cuda()
    x = torch.randn(B, L, 768, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, L, 768, device="cuda").to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        blk.use_fused = fused
        blk.zero_grad(set_to_none=True)
        xg = x.clone().requires_grad_(True)
        y = blk(xg, cs, seq)
        y.backward(dy)
        outs.append((y.detach(), xg.grad, {n: p.grad.clone() for n, p in blk.named_parameters() if p.grad is not None}))
    (y1, dx1, g1), (y2, dx2, g2) = outs
    assert rel_err(y1, y2) <= 1e-2 and rel_err(dx1, dx2) <= 2e-2
    assert set(g1) == set(g2) and len(g1) == 14
    for n in g1:
        assert rel_err(g1[n], g2[n]) <= 2e-2, n


def test_jit_h_like_denoiser_head_dim_80():
    """A JiT-H-shaped model in miniature (head_dim 80, rope_axes_dims [16, 32, 32], ragged SwiGLU width, NF4 + LoRA): the
    fused block path runs, and prediction / loss / LoRA gradients match the fp32 oracle on the same weights."""
    from vision_pt_b200 import ops
    from vision_pt_b200 import train as T
    from vision_pt_b200.jit import DenoiserConfig
    dev = torch.device("cuda")
    cfg = DenoiserConfig(patch_size=16, in_channels=3, out_channels=3, hidden_size=320, depth=2, num_heads=4, mlp_ratio=4.0,
                         bottleneck_dim=32, num_time_tokens=4, rope_axes_dims=[16, 32, 32], context_dim=64,
                         context_start_block=1)
    cfgd = cfg.model_dump()
    net = T.build_jit_qlora(cfg, rank=16, alpha=16.0, device=dev, seed=21, lora_up_std=0.02)
    assert all(b.fused_eligible(torch.empty(1, 1, 320, device=dev, dtype=torch.bfloat16)) for b in net.blocks)
    B, H, W, Tn = 3, 64, 96, 16
    g = torch.Generator().manual_seed(2)
    image = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16)
    t = torch.rand(B, generator=g).to(torch.bfloat16)
    ctx = (torch.randn(B, Tn, 64, generator=g) * 0.5).to(torch.bfloat16)
    mask = (torch.arange(Tn).unsqueeze(0) < torch.tensor([16, 9, 3]).unsqueeze(1)).to(torch.int64)
    size = torch.tensor([[H, W]]).repeat(B, 1)
    clean = torch.randn(B, 3, H, W, generator=g).to(torch.bfloat16)
    pred = net(image=image.to(dev), timestep=t.to(dev), context=ctx.to(dev), original_size=size.to(dev), target_size=size.to(dev),
               crop_coords=torch.zeros_like(size).to(dev), context_mask=mask.to(dev))
    loss = ops.flow_loss(pred, clean.to(dev), loss_target="image")
    loss.backward()
    P = {}
    for name, p in net.state_dict().items():
        if ".weight." not in name:
            P[name] = p.detach().float().cpu()
    for name, mod in net.named_modules():
        qs = getattr(mod, "quant_state", None)
        if qs is not None:
            st = on.Nf4State(packed=qs.packed.cpu(), absmax=qs.absmax.cpu(), nested_absmax=qs.nested_absmax.cpu(),
                             nested_code=qs.nested_code.cpu(), code=qs.code.cpu(), offset=float(qs.offset),
                             shape=tuple(qs.shape), dtype=qs.dtype)
            P[f"{name}.weight"] = on.dequantize_nf4(st).float()
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_down" in k or "lora_up" in k}
    P.update(leaves)
    y = oj.jit_forward(P, cfgd, image.float(), t.float(), ctx.float(), size, size, torch.zeros_like(size), context_mask=mask, alpha=16.0)
    ref_loss = torch.nn.functional.mse_loss(y, clean.float())
    ref_loss.backward()
    assert rel_err(pred, y) <= 3e-2
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * abs(float(ref_loss))
    worst = max(rel_err(p.grad, leaves[n].grad) for n, p in net.named_parameters() if p.requires_grad)
    assert worst <= 5e-2, worst
