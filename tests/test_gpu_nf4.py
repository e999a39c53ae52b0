"""NF4 kernels through the C ABI vs the CPU oracle: bit-exact dequantisation, quantisation, state-dict interchange."""
import pytest
import torch

from oracle import nf4 as on

pytestmark = pytest.mark.gpu


def _to_dev(st: on.Nf4State):
    from vision_pt_b200 import ops
    return ops.Nf4Tensors(st.packed.cuda(), st.absmax.cuda(), st.nested_absmax.cuda(), st.nested_code.cuda(), st.code.cuda(),
                          st.offset, st.shape, st.dtype)


@pytest.mark.parametrize("shape", [(48, 64), (768, 768), (2048, 768), (768, 2048), (1024, 1024), (640, 2048), (96, 200)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_dequant_bit_exact(shape, dtype):
    from vision_pt_b200 import ops
    torch.manual_seed(shape[0] * 7 + shape[1])
    w = (torch.randn(shape) * 0.02).to(dtype)
    st = on.quantize_nf4(w)
    ref = on.dequantize_nf4(st)
    got = ops.nf4_dequantize(_to_dev(st)).cpu()
    assert got.dtype == dtype and torch.equal(got.view(torch.uint8), ref.view(torch.uint8))
    # bf16 target for a weight quantised from another dtype (what MatMul4Bit's .to(x.dtype) sees is a second rounding)
    got_bf = ops.nf4_dequantize(_to_dev(st), torch.bfloat16).cpu()
    st.dtype = torch.bfloat16
    assert torch.equal(got_bf, on.dequantize_nf4(st))


def test_dequant_extreme_statistics():
    """Blocks of zeros, one huge outlier, denormal-ish values: same bits as the oracle."""
    from vision_pt_b200 import ops
    w = torch.zeros(64, 64)
    w[3] = 1e-30
    w[5, 7] = 6e4
    w[9] = torch.linspace(-1, 1, 64)
    st = on.quantize_nf4(w.to(torch.bfloat16))
    assert torch.equal(ops.nf4_dequantize(_to_dev(st)).cpu().view(torch.int16), on.dequantize_nf4(st).view(torch.int16))


@pytest.mark.parametrize("shape,dtype", [((768, 768), torch.bfloat16), ((2048, 768), torch.bfloat16), ((256, 320), torch.float32),
                                         ((128, 64), torch.float16)])
def test_quantize_matches_oracle(shape, dtype):
    from vision_pt_b200 import ops
    from vision_pt_b200.modules.quant import NF4_CODE, nested_code_table
    torch.manual_seed(1)
    w = (torch.randn(shape) * 0.02).to(dtype)
    ref = on.quantize_nf4(w)
    got = ops.nf4_quantize(w.cuda(), nested_code_table(), torch.tensor(NF4_CODE))
    assert torch.equal(got.packed.cpu(), ref.packed)                       # 4-bit codes: identical
    assert torch.equal(got.nested_code.cpu(), ref.nested_code) and torch.equal(got.code.cpu(), ref.code)
    # the statistics depend on mean(absmax), whose fp32 summation order differs: codes may move by one step
    assert abs(got.offset - ref.offset) <= 1e-6 * abs(ref.offset)
    assert (got.absmax.cpu().int() - ref.absmax.int()).abs().max() <= 1
    d_got, d_ref = ops.nf4_dequantize(got).cpu().float(), on.dequantize_nf4(ref).float()
    assert (d_got - d_ref).abs().max() <= 0.02 * d_ref.abs().max()
    # and the kernel dequantises ITS OWN state exactly like the oracle does
    own = on.Nf4State(got.packed.cpu(), got.absmax.cpu(), got.nested_absmax.cpu(), got.nested_code.cpu(), got.code.cpu(),
                      got.offset, got.shape, got.dtype)
    assert torch.equal(ops.nf4_dequantize(got).cpu(), on.dequantize_nf4(own))


def test_linear_module_state_dict_interchange():
    """quantize_inplace -> state_dict has the bitsandbytes key set -> reload as prequantised -> same output
    (reference tests/test_modules_quant.py:153-193)."""
    import torch.nn as nn
    from vision_pt_b200.modules.quant import NF4Linear, quantize_inplace, replace_by_prequantized_weights

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(128, 64)

    torch.manual_seed(0)
    net = Net().to(torch.bfloat16)
    quantize_inplace(net, "bnb_nf4", ["linear"])
    assert isinstance(net.linear, NF4Linear)
    net.cuda()
    sd = net.state_dict()
    assert sd["linear.weight"].dtype == torch.uint8
    for k in ("absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state.bitsandbytes__nf4"):
        assert f"linear.weight.{k}" in sd
    x = torch.randn(5, 128, dtype=torch.bfloat16, device="cuda")
    y1 = net.linear(x)
    net2 = Net().to(torch.bfloat16)
    cpu_sd = {k: v.cpu() for k, v in sd.items()}
    replace_by_prequantized_weights(net2, cpu_sd)
    net2.load_state_dict(cpu_sd, assign=True)
    net2.cuda()
    assert torch.equal(net2.linear(x), y1)
    # oracle reads the same checkpoint
    st = on.Nf4State.from_dict(cpu_sd["linear.weight"], {k[len("linear.weight."):]: v for k, v in cpu_sd.items() if k.startswith("linear.weight.")})
    assert torch.equal(net.linear.dequantize().cpu(), on.dequantize_nf4(st))
    # fp16 in -> fp16 out
    assert net.linear(x.half()).dtype == torch.float16


@pytest.mark.parametrize("shapes", [[(768, 768), (2048, 768), (768, 2048)], [(1024, 2730), (2730, 1024), (192, 341)],
                                    [(1280, 3413), (3413, 1280)]])
@pytest.mark.parametrize("transposed", [False, True])
def test_batched_dequant_slots_bit_exact(shapes, transposed):
    """vpt_nf4_dequant_batch (the per-block launch that feeds the CTA-pair GEMM): every slot holds exactly the oracle's
    dequantised weight -- row-major [N, ld] for the forward, transposed [K, ld] for the backward -- including the ragged
    2730 / 3413 / 341 widths whose 64-blocks and bytes straddle rows; with LoRA given, the transposed slot also carries
    lora_up^T and lora_down^T behind the weight."""
    from vision_pt_b200 import ops
    torch.manual_seed(len(shapes) * 7 + int(transposed))
    qs, refs, downs, ups = [], [], [], []
    for (N, K) in shapes:
        w = (torch.randn(N, K) * 0.05).to(torch.bfloat16)
        st = on.quantize_nf4(w)
        refs.append(on.dequantize_nf4(st))
        qs.append(ops.Nf4Tensors(st.packed.cuda(), st.absmax.cuda(), st.nested_absmax.cuda(), st.nested_code.cuda(), st.code.cuda(),
                                 float(st.offset), (N, K), torch.bfloat16))
        kp = (K + 7) // 8 * 8
        d = torch.zeros(16, kp, dtype=torch.bfloat16)
        d[:, :K] = torch.randn(16, K).to(torch.bfloat16)
        downs.append(d.cuda()[:, :K])
        ups.append(torch.randn(N, 16).to(torch.bfloat16).cuda())
    slots = ops.dequant_block(qs, downs, ups, transposed)
    torch.cuda.synchronize()
    for (N, K), slot, ref, d, u in zip(shapes, slots, refs, downs, ups):
        buf = slot.view(torch.bfloat16)
        if not transposed:
            ld = (K + 7) // 8 * 8
            got = buf[:N * ld].view(N, ld)[:, :K]
            assert torch.equal(got.cpu(), ref), (N, K)
        else:
            ld = (N + 7) // 8 * 8
            got = buf[:K * ld].view(K, ld)[:, :N]
            assert torch.equal(got.cpu(), ref.t()), (N, K)
            upT = buf[K * ld:K * ld + 16 * ld].view(16, ld)[:, :N]
            downT = buf[K * ld + 16 * ld:K * ld + 16 * ld + K * 16].view(K, 16)
            assert torch.equal(upT.cpu(), u.cpu().t()) and torch.equal(downT.cpu(), d.cpu().t()), (N, K)
