"""Generates tests/golden/block_family_vectors.pt by running the REFERENCE's own SDXL and CogView4 transformer blocks (and
the JiT extension pieces: PoPE, U-JiT skip merge) imported from /root/reference -- which only exists in the build
container -- on small seeded inputs in bf16, with LoRA installed by the reference's own PEFT entry point.
Run:  python tests/golden/make_golden_blocks.py

Import recipe: oracle/refimport.py (the package __init__ files that pull accelerate / bitsandbytes are bypassed,
SURVEY.md section 8c); every module that computes is the reference's own file.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refimport  # noqa: E402

refimport.load("/root/reference")

from src.models.cogview4.denoiser import FinalAdaLayerNorm  # noqa: E402
from src.models.cogview4.denoiser import TransformerBlock as CogBlock  # noqa: E402
from src.models.jit.extension.pope import PopeEmbedder, apply_pope  # noqa: E402
from src.models.sdxl.denoiser import TransformerBlock as SdxlBlock  # noqa: E402
from src.modules.peft import PeftTargetConfig  # noqa: E402
from src.modules.peft.lora import LoRAConfig  # noqa: E402

torch.manual_seed(4321)
G = {}
BF = torch.bfloat16


def with_lora(block, keys, rank=16, alpha=8.0):
    block.to(BF).requires_grad_(False)
    PeftTargetConfig(include_keys=keys, config=LoRAConfig(rank=rank, alpha=alpha)).replace_to_peft_layer(block)
    for n, p in block.named_parameters():
        if "lora_up" in n:
            torch.nn.init.normal_(p, std=0.05)
        p.requires_grad_("lora_" in n and "alpha" not in n)
    return block


def run(block, inputs: dict, grad_inputs: list[str], outs_to_tuple=lambda o: (o,)):
    leaves = {k: v.clone().requires_grad_(True) for k, v in inputs.items() if k in grad_inputs}
    call = {**inputs, **leaves}
    outs = outs_to_tuple(block(**call))
    dys = [torch.randn_like(o) for o in outs]
    torch.autograd.backward(list(outs), dys)
    return {"outputs": [o.detach() for o in outs], "d_outputs": dys, "d_inputs": {k: v.grad for k, v in leaves.items()},
            "lora_grads": {n: p.grad.clone() for n, p in block.named_parameters() if p.requires_grad}}


# ---------------------------------------------------------------- SDXL TransformerBlock (src/models/sdxl/denoiser.py:213-280)
blk = SdxlBlock(hidden_dim=128, num_heads=2, head_dim=64, context_dim=96)
for n, p in blk.named_parameters():
    torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, mean=1.0 if "norm" in n and "weight" in n else 0.0, std=0.1)
blk = with_lora(blk, ["attn1", "attn2", ".ff.", "ff."])
state = {k: v.clone() for k, v in blk.state_dict().items()}
inp = {"hidden_states": torch.randn(2, 70, 128).to(BF), "context": torch.randn(2, 11, 96).to(BF), "time_embedding": None}
res = run(blk, inp, ["hidden_states"])
G["sdxl_block"] = {"cfg": dict(hidden_dim=128, num_heads=2, head_dim=64, context_dim=96), "state": state, "inputs": inp, "alpha": 8.0,
                   "rank": 16, **res}

# ---------------------------------------------------------------- CogView4 TransformerBlock (src/models/cogview4/denoiser.py:346-423)
blk = CogBlock(hidden_dim=256, num_attention_heads=4, time_embed_dim=64)
for n, p in blk.named_parameters():
    torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, std=0.1)
blk = with_lora(blk, ["attn1", "ff"])
state = {k: v.clone() for k, v in blk.state_dict().items()}
S, hd = 48, 64
ang = torch.randn(S, hd // 2) * 2.0
freqs = torch.cat([ang, ang], dim=-1)                       # RoPE.forward: cat([freqs, freqs]) (denoiser.py:481)
rot = (freqs.cos(), freqs.sin())
inp = {"hidden_states": torch.randn(2, S, 256).to(BF), "encoder_hidden_states": torch.randn(2, 9, 256).to(BF),
       "time_embed": torch.randn(2, 64).to(BF), "image_rotary_emb": rot}
res = run(blk, inp, ["hidden_states", "encoder_hidden_states", "time_embed"], outs_to_tuple=lambda o: tuple(o))
G["cogview4_block"] = {"cfg": dict(hidden_dim=256, num_attention_heads=4, time_embed_dim=64), "state": state, "inputs": inp,
                       "alpha": 8.0, "rank": 16, **res}

fin = FinalAdaLayerNorm(hidden_dim=256, condition_dim=64)
for p in fin.parameters():
    torch.nn.init.normal_(p, std=0.1)
fin.to(BF)
x, cond = torch.randn(2, S, 256).to(BF), torch.randn(2, 64).to(BF)
G["cogview4_final_norm"] = {"state": {k: v.clone() for k, v in fin.state_dict().items()}, "x": x, "cond": cond, "y": fin(x, cond).detach()}

# ---------------------------------------------------------------- PoPE (src/models/jit/extension/pope.py:6-38)
emb = PopeEmbedder(pope_theta=256.0, axes_dims=[16, 24, 24], axes_lens=[256, 128, 128])
pos = torch.cat([emb.prepare_image_position_ids(32, 48, 16, 3), emb.prepare_context_position_ids(5, 0)], dim=0)   # [L, 3]
fc = emb(pos.unsqueeze(0))                                     # [1, L, 64] complex64
L = fc.shape[1]
xq = torch.randn(2, 3, L, 64).to(BF)
bias = torch.randn(3, 64) * 0.5
leaf = xq.clone().requires_grad_(True)
y = apply_pope(leaf, fc.repeat(2, 1, 1), bias)
dy = torch.randn_like(y)
y.backward(dy)
G["pope"] = {"x": xq, "freqs_cis": fc[0].clone(), "bias": bias, "y": y.detach(), "dy": dy, "dx": leaf.grad.clone(),
             "y_nobias": apply_pope(xq, fc.repeat(2, 1, 1), None).detach()}

# ---------------------------------------------------------------- U-JiT block with a long skip (extension/uvit.py:28-147)
from src.models.jit.denoiser import RopeEmbedder  # noqa: E402
from src.models.jit.extension.cross import CrossJiTBlock  # noqa: E402
from src.models.jit.extension.uvit import UJiTBlock  # noqa: E402

rope = RopeEmbedder(rope_theta=256.0, axes_dims=[16, 24, 24], axes_lens=[256, 128, 128], zero_centered=[False, True, True])
img_freqs = rope(rope.prepare_image_position_ids(32, 48, 16, 3).unsqueeze(0))          # [1, 6, 32] complex
ctx_freqs = rope(rope.prepare_context_position_ids(7, 0).unsqueeze(0))                 # [1, 7, 32]


def init_block(blk):
    for n, p in blk.named_parameters():
        torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, mean=1.0 if "norm" in n else 0.0, std=0.1)
    return blk


blk = with_lora(init_block(UJiTBlock(hidden_dim=128, num_heads=2, has_skip_connection=True)), ["attn.", "mlp.", "skip_merge"])
state = {k: v.clone() for k, v in blk.state_dict().items()}
freqs = torch.cat([img_freqs, ctx_freqs], dim=1)                                       # [1, 13, 32]
L = freqs.shape[1]
mask = torch.ones(2, L)
mask[1, L - 3:] = 0
inp = {"hidden_states": torch.randn(2, L, 128).to(BF), "rope_freqs": freqs.repeat(2, 1, 1),
       "skip_hidden_states": torch.randn(2, L, 128).to(BF), "mask": mask}
res = run(blk, inp, ["hidden_states", "skip_hidden_states"])
G["ujit_block"] = {"state": state, "inputs": inp, "alpha": 8.0, "rank": 16, **res}

# ---------------------------------------------------------------- cross-attention JiT block (extension/cross.py:281-385)
blk = with_lora(init_block(CrossJiTBlock(hidden_dim=128, num_heads=2)), ["attn.", "mlp."])
state = {k: v.clone() for k, v in blk.state_dict().items()}
cmask = torch.ones(2, 7)
cmask[0, 4:] = 0
inp = {"image_hidden_states": torch.randn(2, 6, 128).to(BF), "context_hidden_states": torch.randn(2, 7, 128).to(BF),
       "image_rope_freqs": img_freqs.repeat(2, 1, 1), "context_rope_freqs": ctx_freqs.repeat(2, 1, 1),
       "image_mask": torch.ones(2, 6), "context_mask": cmask}
res = run(blk, inp, ["image_hidden_states", "context_hidden_states"], outs_to_tuple=lambda o: (o[0],))
G["cross_jit_block"] = {"state": state, "inputs": inp, "alpha": 8.0, "rank": 16, **res}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "block_family_vectors.pt")
torch.save(G, out)
print("wrote", out, {k: list(v.keys())[:6] for k, v in G.items()})
