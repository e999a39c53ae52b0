"""Host-side mirror of the reference API that needs no GPU: key selection, module swapping, key names, errors."""
import pytest
import torch
import torch.nn as nn


def test_get_target_keys_reference_lists():
    """reference tests/test_utils.py:12-50 semantics: substring include, regex include, exclude wins."""
    from vision_pt_b200.modules.state_dict import RegexMatch, get_target_keys
    keys = ["model.layer1.attn.to_q", "model.layer1.attn.to_k", "model.layer1.mlp.fc", "model.layer2.attn.to_q", "head"]
    assert sorted(get_target_keys(["attn"], [], keys)) == ["model.layer1.attn.to_k", "model.layer1.attn.to_q", "model.layer2.attn.to_q"]
    assert sorted(get_target_keys(["attn"], ["to_k"], keys)) == ["model.layer1.attn.to_q", "model.layer2.attn.to_q"]
    assert sorted(get_target_keys([RegexMatch(regex=r"model\.layer1\..*")], [RegexMatch(regex=r".*\.mlp\..*")], keys)) == \
        ["model.layer1.attn.to_k", "model.layer1.attn.to_q"]
    assert get_target_keys([], [], keys) == []


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.layer1 = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 64))
        self.head = nn.Linear(64, 8)


def test_peft_replacement_and_key_names():
    """reference tests/test_peft.py:67-127: which layers are wrapped, what trains, adapter key names."""
    from vision_pt_b200.modules.peft import LoRAConfig, LoRALinear, PeftTargetConfig, get_adapter_parameters
    m = _Toy()
    m.requires_grad_(False)
    PeftTargetConfig(include_keys=["layer1"], exclude_keys=["layer1.2"], config=LoRAConfig(rank=4, alpha=2.0)).replace_to_peft_layer(m)
    assert isinstance(m.layer1[0], LoRALinear) and isinstance(m.layer1[2], nn.Linear) and isinstance(m.head, nn.Linear)
    lay = m.layer1[0]
    assert lay.lora_down.weight.shape == (4, 64) and lay.lora_up.weight.shape == (128, 4)
    assert torch.count_nonzero(lay.lora_up.weight) == 0 and torch.count_nonzero(lay.lora_down.weight) > 0
    lay.requires_grad_(True)
    trainable = sorted(n for n, p in m.named_parameters() if p.requires_grad)
    assert trainable == ["layer1.0.lora_down.weight", "layer1.0.lora_up.weight"]
    assert sorted(get_adapter_parameters(m)) == ["layer1.0.alpha", "layer1.0.lora_down.weight", "layer1.0.lora_up.weight"]
    m.train()
    assert not lay.linear.training
    assert abs(lay.scale - 0.5) < 1e-12
    with pytest.raises(ValueError):
        PeftTargetConfig(include_keys=[], config=LoRAConfig(rank=4))


def test_peft_weight_round_trip():
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig, get_adapter_parameters, load_peft_weight
    a, b = _Toy(), _Toy()
    PeftTargetConfig(include_keys=["head"], config=LoRAConfig(rank=8, alpha=3.0)).replace_to_peft_layer(a)
    nn.init.normal_(a.head.lora_up.weight)
    sd = {k: v.clone() for k, v in get_adapter_parameters(a).items()}
    load_peft_weight(b, sd)
    assert torch.equal(b.head.lora_up.weight, a.head.lora_up.weight) and b.head.rank == 8 and abs(b.head.scale - 3.0 / 8) < 1e-6
    with pytest.raises(ValueError):
        load_peft_weight(_Toy(), {"x": torch.zeros(1)})


def test_quant_registry_swaps_and_errors():
    """reference tests/test_modules_quant.py:21-126 for the bnb_nf4 type; the other types are declared out of scope."""
    from vision_pt_b200.modules.quant import (NF4Linear, get_quant_type_from_children_dict, replace_by_prequantized_weights,
                                              replace_to_quant_linear, validate_quant_type)
    m = _Toy()
    replace_to_quant_linear(m, "bnb_nf4", ["layer1"])
    assert isinstance(m.layer1[0], NF4Linear) and isinstance(m.layer1[2], NF4Linear) and not isinstance(m.head, NF4Linear)
    assert isinstance(m.layer1[0], nn.Linear) and m.layer1[0].quant_type == "nf4"
    assert m.layer1[0].weight.is_meta and not m.layer1[0].weight.requires_grad
    with pytest.raises(ValueError):
        validate_quant_type("int3")
    with pytest.raises(NotImplementedError):
        replace_to_quant_linear(_Toy(), "bnb_int8", ["head"])
    sd = {"head.weight": torch.zeros(256, 1, dtype=torch.uint8), "head.weight.absmax": torch.zeros(8, dtype=torch.uint8),
          "head.weight.quant_state.bitsandbytes__nf4": torch.zeros(4, dtype=torch.uint8), "layer1.0.weight": torch.zeros(128, 64)}
    m2 = _Toy()
    replace_by_prequantized_weights(m2, sd)
    assert isinstance(m2.head, NF4Linear) and not isinstance(m2.layer1[0], NF4Linear)
    assert get_quant_type_from_children_dict({"quant_state.bitsandbytes__nf4": torch.zeros(1)}) == "bnb_nf4"
    with pytest.raises(ValueError):
        get_quant_type_from_children_dict({"foo": torch.zeros(1)})
    with pytest.raises(RuntimeError):
        m.layer1[0](torch.zeros(2, 64))          # no quantised weight yet -> loud failure, never a silent fallback


def test_prequantized_state_dict_loads_on_cpu():
    from oracle import nf4 as on
    from vision_pt_b200.modules.quant import NF4Linear
    w = (torch.randn(64, 128) * 0.05).to(torch.bfloat16)
    st = on.quantize_nf4(w)
    sd = {"weight": st.packed, "bias": torch.zeros(64, dtype=torch.bfloat16)}
    sd.update({f"weight.{k}": v for k, v in st.as_dict().items()})
    lin = NF4Linear(128, 64)
    lin.load_state_dict(sd, assign=True)
    assert lin.is_quantized and lin.quant_state.shape == (64, 128) and lin.quant_state.dtype == torch.bfloat16
    out = lin.state_dict()
    assert set(out) == set(sd)
    for k in sd:
        assert torch.equal(out[k], sd[k]), k


def test_cpu_tensors_are_rejected():
    from vision_pt_b200 import ops
    with pytest.raises(RuntimeError):
        ops.rms_norm(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        ops.attention(*(torch.zeros(1, 1, 8, 64, dtype=torch.bfloat16) for _ in range(3)))


def test_key_lengths_from_mask():
    from vision_pt_b200.modules.attention import key_lengths_from_mask
    km = torch.tensor([[1, 1, 1, 0], [1, 1, 0, 0]], dtype=torch.bool)
    assert key_lengths_from_mask(km, 2, 4).tolist() == [3, 2]
    assert key_lengths_from_mask(km.view(2, 1, 1, 4).expand(-1, 3, 5, -1), 2, 4).tolist() == [3, 2]
    assert key_lengths_from_mask(None, 2, 4) is None
    with pytest.raises(NotImplementedError):
        key_lengths_from_mask(torch.ones(2, 3, 5, 4, dtype=torch.bool), 2, 4)


def test_jit_config_and_model_structure():
    """Parameter names are the reference's (checkpoint / PEFT key compatibility): compare with the golden state dict."""
    import os
    from vision_pt_b200.jit import Denoiser, DenoiserConfig, JiT_B_16_Config
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.pt"), weights_only=False)["denoiser_f32"]
    m = Denoiser(DenoiserConfig(**g["cfg"]))
    assert set(m.state_dict()) == set(g["state"])
    for k, v in m.state_dict().items():
        assert v.shape == g["state"][k].shape, k
    b = JiT_B_16_Config()
    assert (b.hidden_size, b.depth, b.num_heads, b.context_start_block) == (768, 12, 12, 4)
    with pytest.raises(AssertionError):
        Denoiser(DenoiserConfig(hidden_size=128, num_heads=4, depth=1))   # rope dims must sum to head_dim


def test_aspect_ratio_buckets_follow_the_reference_rule():
    """tools/bench_arb.py restates generate_buckets (src/dataset/aspect_ratio_bucket.py:20-60): base 512 / step 64 / min 256
    gives the 9 buckets BASELINE.json configs[3] trains on; the reference's defaults (1024 / 64 / 384) give 21."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bench_arb.py")
    spec = importlib.util.spec_from_file_location("bench_arb", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.buckets(512, 64, 256)
    assert sorted(got) == sorted([(512, 512), (448, 576), (576, 448), (384, 704), (704, 384), (320, 832), (832, 320),
                                  (256, 1024), (1024, 256)])
    big = mod.buckets(1024, 64, 384)
    assert (1024, 1024) in big and (384, 2752) in big and (2752, 384) in big and all(h % 64 == 0 and w % 64 == 0 for h, w in big)
    assert len(big) == len(set(big)) == 21


def test_padded_weight_copy_lives_on_the_weight_object():
    """ops.padded_weight: a ragged [N, K] weight gets a row-padded copy that is cached ON the tensor (a cache keyed by
    address once handed a freed weight's copy to the next tensor allocated there) and is rebuilt after in-place updates."""
    import torch

    from vision_pt_b200 import ops
    a = torch.randn(8, 13).to(torch.bfloat16)
    pa = ops.padded_weight(a)
    assert pa.shape == (8, 13) and pa.stride(0) == 16 and torch.equal(pa, a) and ops.padded_weight(a) is pa
    b = torch.randn(8, 13).to(torch.bfloat16)
    assert torch.equal(ops.padded_weight(b), b) and not torch.equal(ops.padded_weight(b), pa)
    a.mul_(2)
    assert torch.equal(ops.padded_weight(a), a)                    # stale copy replaced after the in-place change
    c = torch.randn(8, 16).to(torch.bfloat16)
    assert ops.padded_weight(c) is c                               # already aligned: no copy


def test_attention_masks_must_be_prefix_key_padding_masks():
    """Advisor finding r1: a non-prefix mask or an additive float mask must be refused, not computed as a prefix."""
    from vision_pt_b200.modules.attention import key_lengths_from_mask, prefix_key_lengths
    ok = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [0, 0, 0, 0, 0]])
    assert prefix_key_lengths(ok).tolist() == [3, 5, 0] and prefix_key_lengths(ok.bool()).dtype == torch.int32
    with pytest.raises(NotImplementedError):
        prefix_key_lengths(torch.tensor([[0, 1, 1, 1, 1]]))            # left padding
    with pytest.raises(NotImplementedError):
        prefix_key_lengths(torch.tensor([[1, 0, 1, 1, 0]]))            # a hole
    with pytest.raises(TypeError):
        prefix_key_lengths(torch.zeros(2, 5))                          # F.sdpa-style additive mask: 0.0 means "attend"
    m4 = ok.bool().view(3, 1, 1, 5).expand(3, 4, 7, 5)
    assert key_lengths_from_mask(m4, 3, 5).tolist() == [3, 5, 0]
    with pytest.raises(NotImplementedError):
        key_lengths_from_mask(torch.ones(3, 4, 7, 5, dtype=torch.bool), 3, 5)   # depends on head / query: not a key-padding mask


def test_lora_alpha_follows_load_state_dict_and_assignment():
    """Advisor finding r1: the kernels scale by a host copy of alpha; it must follow the stored parameter."""
    from vision_pt_b200.modules.peft import LoRAConfig, LoRALinear
    lay = LoRALinear(LoRAConfig(rank=4, alpha=2.0, dtype="float32"), nn.Linear(16, 8))
    assert lay.scale == 0.5
    sd = {k: v.clone() for k, v in lay.state_dict().items()}
    sd["alpha"] = torch.tensor(6.0)
    lay.load_state_dict(sd)
    assert float(lay.alpha) == 6.0 and lay.scale == 1.5
    lay.alpha = nn.Parameter(torch.tensor(1.0), requires_grad=False)
    assert lay.scale == 0.25
    wrapper = nn.Sequential(lay)                                     # through a parent module's load_state_dict as well
    sd = {k: v.clone() for k, v in wrapper.state_dict().items()}
    sd["0.alpha"] = torch.tensor(8.0)
    wrapper.load_state_dict(sd)
    assert lay.scale == 2.0


def test_gradient_chunk_plan_follows_block_order():
    """The chunks of the overlapped all-reduce: contiguous flat ranges, last blocks first, covering the buffer once."""
    from types import SimpleNamespace

    from vision_pt_b200.jit import Denoiser, DenoiserConfig
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig
    from vision_pt_b200.train import LORA_TARGET, FlatLoRA, JiTQLoRATrainStep
    cfg = DenoiserConfig(patch_size=16, hidden_size=128, depth=6, num_heads=2, bottleneck_dim=32, context_dim=64)
    net = Denoiser(cfg).to(torch.bfloat16)
    net.requires_grad_(False)
    PeftTargetConfig(include_keys=[LORA_TARGET], config=LoRAConfig(rank=16, alpha=16.0)).replace_to_peft_layer(net)
    for n, p in net.named_parameters():
        p.requires_grad_(".lora_" in n)
    flat = FlatLoRA(net)
    stub = SimpleNamespace(model=net, flat=flat)
    for n_chunks in (1, 2, 3):
        plan = JiTQLoRATrainStep._chunk_plan(stub, n_chunks)
        assert len(plan) == n_chunks
        assert plan[0][2] == flat.numel and plan[-1][1] == 0 and plan[-1][0] == 0
        for (fb_a, lo_a, hi_a), (fb_b, lo_b, hi_b) in zip(plan, plan[1:]):
            assert lo_a == hi_b and fb_a > fb_b                       # adjacent, and in the order backward finishes them
        names = {id(p): n for n, p in net.named_parameters()}
        for fb, lo, hi in plan:                                       # a chunk holds exactly the blocks >= its first block
            inside = [int(names[id(p)].split(".")[1]) for p, off in zip(flat.params, flat.offsets) if lo <= off < hi]
            assert min(inside) == fb
    assert JiTQLoRATrainStep._chunk_plan(stub, 4) == [(0, 0, flat.numel)]   # fewer than 2 blocks per chunk: one exchange


def test_flat_lora_keeps_qkv_adapters_adjacent_and_stacked_views_are_zero_copy():
    """train.FlatLoRA lays the q / k / v lora_down matrices (then the three lora_up matrices) back to back, so the fused
    block's single q | k | v GEMM takes them as two views of the flat buffer (ops.stacked); unrelated tensors are copied."""
    from vision_pt_b200 import ops
    from vision_pt_b200.jit import Denoiser, DenoiserConfig
    from vision_pt_b200.modules.peft import LoRAConfig, PeftTargetConfig
    from vision_pt_b200.train import LORA_TARGET, FlatLoRA
    cfg = DenoiserConfig(patch_size=16, hidden_size=128, depth=2, num_heads=2, bottleneck_dim=32, context_dim=64)
    net = Denoiser(cfg).to(torch.bfloat16)
    net.requires_grad_(False)
    PeftTargetConfig(include_keys=[LORA_TARGET], config=LoRAConfig(rank=16, alpha=16.0)).replace_to_peft_layer(net)
    for n, p in net.named_parameters():
        p.requires_grad_(".lora_" in n)
    before = {n: p.detach().clone() for n, p in net.named_parameters() if ".lora_" in n}
    flat = FlatLoRA(net)
    for n, p in net.named_parameters():                      # re-pointing the parameters into the flat buffer keeps their values
        if ".lora_" in n:
            assert torch.equal(p.detach(), before[n]), n
    attn = net.blocks[1].attn
    downs = [m.lora_down.weight for m in (attn.to_q, attn.to_k, attn.to_v)]
    ups = [m.lora_up.weight for m in (attn.to_q, attn.to_k, attn.to_v)]
    d, u = ops.stacked([t.detach() for t in downs]), ops.stacked([t.detach() for t in ups])
    assert d.shape == (48, 128) and u.shape == (384, 16)
    assert d.data_ptr() == downs[0].data_ptr() and u.data_ptr() == ups[0].data_ptr()          # views, not copies
    assert torch.equal(d, torch.cat([t.detach() for t in downs])) and torch.equal(u, torch.cat([t.detach() for t in ups]))
    mixed = ops.stacked([downs[0].detach(), downs[2].detach()])                                 # not adjacent: a copy
    assert mixed.data_ptr() != downs[0].data_ptr() and torch.equal(mixed, torch.cat([downs[0].detach(), downs[2].detach()]))
    # every trainable matrix is in the buffer exactly once, slices do not overlap
    spans = sorted((off, off + p.numel()) for p, off in zip(flat.params, flat.offsets))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and len(spans) == 2 * 7 * cfg.depth


def test_fused_rows_view_needs_gap_free_slots():
    from vision_pt_b200 import ops
    n, k = 128, 64
    need = n * k * 2
    arena = torch.zeros(3 * need + 512, dtype=torch.uint8)
    slots = [arena[i * need:(i + 1) * need] for i in range(3)]
    v = ops.fused_rows_view(slots, n, k)
    assert v is not None and v.shape == (3 * n, k) and v.dtype == torch.bfloat16 and v.data_ptr() == arena.data_ptr()
    v[n, 0] = 1.0                                              # row n of the view is row 0 of the second slot
    assert slots[1].view(torch.bfloat16)[0] == 1.0
    gapped = [arena[0:need], arena[need + 256:2 * need + 256]]
    assert ops.fused_rows_view(gapped, n, k) is None


def test_transposed_weight_copy_is_cached_on_the_weight_and_padded():
    """ops.transposed_weight: the [K, N] operand the dense frozen linears' input gradient runs on (rows padded to 8 elements,
    zeros in the padding), cached on the weight object and rebuilt after an in-place modification."""
    from vision_pt_b200 import ops
    w = torch.randn(5, 13).to(torch.bfloat16)                # N = 5 (ragged), K = 13
    tw = ops.transposed_weight(w)
    assert tw.shape == (13, 5) and tw.stride() == (8, 1) and torch.equal(tw, w.t())
    assert ops.transposed_weight(w) is tw
    w.mul_(2)
    tw2 = ops.transposed_weight(w)
    assert tw2 is not tw and torch.equal(tw2, w.t())
    # through the padded view of a ragged weight: the cache sits on that view, which the weight keeps alive
    pw = ops.padded_weight(w)
    assert ops.transposed_weight(pw) is ops.transposed_weight(ops.padded_weight(w))


def test_token_prefix_and_packed_tokens_without_a_device():
    """The slice helpers fall back to torch for tensors the row-copy kernel does not take (here: CPU tensors) with the same
    values and gradients; noise_mix, like every kernel wrapper, refuses CPU tensors instead of computing on the host."""
    from vision_pt_b200 import ops
    x = torch.randn(2, 7, 6, requires_grad=True)
    y = ops.token_prefix(x, 4)
    assert torch.equal(y, x[:, :4]) and y.is_contiguous()
    y.sum().backward()
    want = torch.zeros(2, 7, 6)
    want[:, :4] = 1
    assert torch.equal(x.grad, want)
    assert ops.token_prefix(x, 7) is x
    v = x.detach()[:, 2:5]
    assert torch.equal(ops.packed_tokens(v), v) and ops.packed_tokens(v).is_contiguous()
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.noise_mix(torch.zeros(2, 3, 4, 4), torch.zeros(2, 3, 4, 4), torch.zeros(2))
