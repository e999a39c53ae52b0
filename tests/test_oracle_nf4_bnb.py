"""oracle/nf4.py against bitsandbytes' own quantize_4bit / dequantize_4bit outputs -- runs as soon as
tests/golden/bnb_nf4_vectors.pt exists (tests/golden/make_golden_bnb.py writes it on a box that has the wheel; neither this
image nor the GPU box does, so today this test is skipped and the NF4 arithmetic stays "parity unpinned")."""
import os

import pytest
import torch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bnb_nf4_vectors.pt")


@pytest.mark.skipif(not os.path.exists(PATH), reason="no bitsandbytes golden vectors (the wheel is not installable here)")
def test_oracle_reproduces_bitsandbytes_bit_for_bit():
    from oracle import nf4 as on
    vec = torch.load(PATH, map_location="cpu", weights_only=False)
    for name, v in vec.items():
        st = on.quantize_nf4(v["weight"])
        assert torch.equal(st.packed.reshape(-1), v["packed"].reshape(-1)), f"{name}: packed codes differ"
        ref = on.Nf4State.from_dict(v["packed"], v["stats"])
        assert torch.equal(st.absmax, ref.absmax) and torch.equal(st.nested_absmax, ref.nested_absmax), f"{name}: statistics differ"
        assert st.offset == ref.offset
        deq = on.dequantize_nf4(ref)
        assert torch.equal(deq.reshape(-1)[::97], v["dequant_sample"]) and torch.equal(deq[:4], v["dequant_first_rows"]), name
        assert float(deq.float().sum()) == v["dequant_sum"]
