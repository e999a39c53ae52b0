"""The oracle (oracle/*.py) against vectors produced by the reference's own modules (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import jit as oj
from oracle import nf4 as on
from tests.helpers import lora_param_dict


def _lora_case(g):
    st = g["state"]
    return oj.lora_linear(g["x"], st["linear.weight"], st["linear.bias"], st["lora_down.weight"], st["lora_up.weight"], g["alpha"])


def test_lora_linear_matches_reference(golden):
    for tag in ("lora_bf16_r16", "lora_bf16_r4", "lora_f32_r8"):
        g = golden[tag]
        assert torch.equal(_lora_case(g), g["y"]), tag


def test_norms_match_reference(golden):
    g = golden["rmsnorm"]
    assert torch.equal(oj.rms_norm_fp32(g["x"], g["w"]), g["y"])
    g = golden["layernorm"]
    assert torch.equal(oj.layer_norm_fp32(g["x"]), g["y"])
    g = golden["adaln"]
    assert torch.equal(oj.adaln_modulate(g["x"], g["scale"], g["shift"], eps=1e-6), g["y"])
    assert torch.equal(oj.gate_residual(g["x"], g["y"], g["gate"]), g["gated"])


def test_rope_matches_reference(golden):
    g = golden["rope"]
    f = oj.rope_freqs_cis(g["cfg"], g["height"], g["width"], g["ctx"])
    assert f.shape == g["freqs_cis"].shape and torch.equal(torch.view_as_real(f), torch.view_as_real(g["freqs_cis"]))
    assert torch.equal(oj.apply_rope(g["x"], f), g["y"])


def test_attention_matches_reference(golden):
    g = golden["attention"]
    assert torch.equal(oj.attention(g["q"], g["k"], g["v"], g["key_mask"]), g["y"])
    # the explicit fp32 form agrees on valid rows within bf16 resolution
    ex = oj.attention_explicit(g["q"], g["k"], g["v"], g["key_mask"].sum(1).long())
    assert (ex - g["y"].float()).abs().max() < 2e-2


def test_block_matches_reference(golden):
    f = oj.rope_freqs_cis(golden["rope"]["cfg"], golden["rope"]["height"], golden["rope"]["width"], golden["rope"]["ctx"])
    for tag, tol in (("block_f32", 1e-5), ("block_bf16", 0.0)):
        g = golden[tag]
        y = oj.jit_block(g["state"], "", g["x"], f, g["key_mask"], num_heads=2)
        assert (y.float() - g["y"].float()).abs().max() <= tol, tag


def test_denoiser_matches_reference(golden):
    g = golden["denoiser_f32"]
    y = oj.jit_forward(g["state"], g["cfg"], **g["inputs"])
    assert (y - g["y"]).abs().max() < 1e-5


def test_denoiser_lora_loss_and_grads_match_reference(golden):
    g, base = golden["denoiser_lora_f32"], golden["denoiser_f32"]["state"]
    P = lora_param_dict(base, g["lora_state"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_" in k}
    P.update(leaves)
    y = oj.jit_forward(P, g["cfg"], **g["inputs"], alpha=g["alpha"])
    assert (y - g["y"]).abs().max() < 1e-5
    loss = oj.velocity_loss(y, g["clean"], g["inputs"]["image"], g["inputs"]["timestep"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    for k, ref in g["grads"].items():
        assert (leaves[k].grad - ref).abs().max() <= 1e-4 * ref.abs().max().clamp_min(1e-8), k


def test_patchify_matches_reference(golden):
    g = golden["patchify"]
    assert torch.equal(oj.patchify(g["image"], 16, order=0), g["patches"])
    assert torch.equal(oj.unpatchify(g["patches"], 3, 32, 48, 16, order=0), g["back"])
    assert torch.equal(g["back"], g["image"])                      # reference tests/test_patch.py identity
    j = golden["jit_unpatchify"]
    assert torch.equal(oj.unpatchify(j["patches"], 3, 32, 48, 16, order=1), j["image"])
    assert torch.equal(oj.patchify(j["image"], 16, order=1), j["patches"])


def test_timestep_embedding_matches_reference(golden):
    g = golden["timestep_embedding"]
    assert torch.allclose(oj.timestep_embedding(g["t"]), g["y"], atol=1e-6)


# ------------------------------------------------------------------------------------------------ NF4 (parity unpinned)
def test_nf4_code_is_the_qlora_construction():
    """bitsandbytes create_normal_map(offset=0.9677083): fp32 linspace -> norm.ppf -> sort -> normalise by the max."""
    from scipy.stats import norm
    offset = 0.9677083
    v1 = norm.ppf(torch.linspace(offset, 0.5, 9)[:-1]).tolist()
    v3 = (-norm.ppf(torch.linspace(offset, 0.5, 8)[:-1])).tolist()
    values = torch.Tensor(v1 + [0] + v3).sort().values
    values /= values.max()
    assert np.array_equal(values.numpy(), on.NF4_CODE)


def test_nf4_dynamic_map_shape():
    code = on.dynamic_map_signed8()
    assert code.shape == (256,) and np.all(np.diff(code) >= 0)
    assert np.isclose(code[0], -0.99296874) and code[-1] == 1.0 and (code == 0).sum() == 1
    assert np.isclose(code[126], -5.5e-07, rtol=1e-3) and code[127] == 0.0 and np.isclose(code[128], 5.5e-07, rtol=1e-3)


def test_nf4_quantize_dequantize_roundtrip():
    torch.manual_seed(0)
    for shape, dt in (((48, 64), torch.bfloat16), ((96, 128), torch.float32), ((768, 768), torch.bfloat16)):
        w = (torch.randn(shape) * 0.02).to(dt)
        st = on.quantize_nf4(w)
        assert st.packed.shape == (w.numel() // 2, 1) and st.packed.dtype == torch.uint8
        assert st.absmax.dtype == torch.uint8 and st.absmax.numel() == w.numel() // 64
        d = on.dequantize_nf4(st)
        assert d.dtype == dt and d.shape == w.shape
        blockmax = w.float().reshape(-1, 64).abs().max(dim=1).values.repeat_interleave(64).reshape(shape)
        # worst-case NF4 step is ~0.152 of the block absmax (gap 1.0 - 0.723 halved) plus the statistics' own error
        assert ((d.float() - w.float()).abs() <= 0.16 * blockmax + 1e-6).all()
        # re-quantising the dequantised tensor is a fixed point of the codes
        st2 = on.quantize_nf4(d)
        assert (st2.packed != st.packed).float().mean() < 0.02


def test_nf4_state_dict_key_set():
    st = on.quantize_nf4((torch.randn(16, 64) * 0.1).to(torch.bfloat16))
    d = st.as_dict()
    assert set(d) == {"absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state.bitsandbytes__nf4"}
    back = on.Nf4State.from_dict(st.packed, d)
    assert torch.equal(on.dequantize_nf4(back), on.dequantize_nf4(st))


# ---------------------------------------------------------------------------------------------- other block families
def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def test_sdxl_block_oracle_matches_reference(golden_blocks):
    """oracle/blocks.py:sdxl_block against the reference's TransformerBlock run live (bf16 on the CPU, LoRA rank 16)."""
    from oracle import blocks as ob
    g = golden_blocks["sdxl_block"]
    P = {k: v.clone() for k, v in g["state"].items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_down" in k or "lora_up" in k}
    P.update(leaves)
    x = g["inputs"]["hidden_states"].clone().requires_grad_(True)
    y = ob.sdxl_block(P, x, g["inputs"]["context"], g["cfg"]["num_heads"], alpha=g["alpha"])
    assert _rel(y, g["outputs"][0]) <= 1e-2                    # same ops in the same dtype: identical up to kernel choice
    y.backward(g["d_outputs"][0])
    assert _rel(x.grad, g["d_inputs"]["hidden_states"]) <= 2e-2
    for n, ref in g["lora_grads"].items():
        assert _rel(leaves[n].grad, ref) <= 2e-2, n


def test_cogview4_block_oracle_matches_reference(golden_blocks):
    from oracle import blocks as ob
    g = golden_blocks["cogview4_block"]
    P = {k: v.clone() for k, v in g["state"].items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items() if "lora_down" in k or "lora_up" in k}
    P.update(leaves)
    inp = g["inputs"]
    x = inp["hidden_states"].clone().requires_grad_(True)
    e = inp["encoder_hidden_states"].clone().requires_grad_(True)
    t = inp["time_embed"].clone().requires_grad_(True)
    y, ye = ob.cogview4_block(P, x, e, t, inp["image_rotary_emb"], g["cfg"]["num_attention_heads"], alpha=g["alpha"])
    assert _rel(y, g["outputs"][0]) <= 1e-2 and _rel(ye, g["outputs"][1]) <= 1e-2
    torch.autograd.backward([y, ye], g["d_outputs"])
    assert _rel(x.grad, g["d_inputs"]["hidden_states"]) <= 2e-2
    assert _rel(e.grad, g["d_inputs"]["encoder_hidden_states"]) <= 2e-2
    assert _rel(t.grad, g["d_inputs"]["time_embed"]) <= 2e-2
    for n, ref in g["lora_grads"].items():
        assert _rel(leaves[n].grad, ref) <= 2e-2, n
    f = golden_blocks["cogview4_final_norm"]
    assert _rel(ob.cogview4_final_norm(f["state"], f["x"], f["cond"]), f["y"]) <= 1e-2


def test_pope_and_tread_oracle(golden_blocks):
    from oracle import blocks as ob
    g = golden_blocks["pope"]
    x = g["x"].clone().requires_grad_(True)
    y = ob.pope(x, g["freqs_cis"], g["bias"])
    assert torch.equal(y, g["y"]) and torch.equal(ob.pope(g["x"], g["freqs_cis"], None), g["y_nobias"])
    y.backward(g["dy"])
    assert _rel(x.grad, g["dx"]) <= 1e-6
    # TREAD: split by a permutation and re-insert = identity (reference class_to_image_tread.py:73-118 and its re-merge)
    t = torch.randn(2, 13, 8)
    perm = torch.randperm(13)
    keep, route = ob.tread_split(t, perm, 5)
    assert keep.shape[1] == 5 and route.shape[1] == 8 and torch.equal(ob.tread_merge(keep, route, perm), t)
