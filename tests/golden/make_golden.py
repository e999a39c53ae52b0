"""Generates tests/golden/reference_vectors.pt by running the REFERENCE's own modules (imported from /root/reference,
which only exists in the build container) on small seeded inputs.  Run:  python tests/golden/make_golden.py

The package __init__ files of src.models.jit pull accelerate / bitsandbytes (absent here), so they are bypassed by
pre-registering bare namespace modules; every module that computes is the reference's own file.
"""
import os
import sys
import types

import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
for name, path in (("src.models", f"{REF}/src/models"), ("src.models.jit", f"{REF}/src/models/jit"),
                   ("src.models.jit.extension", f"{REF}/src/models/jit/extension")):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m

from src.models.jit.config import DenoiserConfig  # noqa: E402
from src.models.jit.denoiser import Denoiser, JiTBlock, RopeEmbedder, apply_rope  # noqa: E402
from src.modules.attention import scaled_dot_product_attention  # noqa: E402
from src.modules.norm import FP32LayerNorm, FP32RMSNorm, SingleAdaLayerNormZero  # noqa: E402
from src.modules.patch import patchify, unpatchify  # noqa: E402
from src.modules.peft import PeftTargetConfig  # noqa: E402
from src.modules.peft.lora import LoRAConfig, LoRALinear  # noqa: E402
from src.modules.timestep.embedding import get_timestep_embedding  # noqa: E402
from src.utils.state_dict import RegexMatch  # noqa: E402

torch.manual_seed(1234)
G = {}

# --- LoRALinear (bf16 and fp32), non-zero lora_up so the branch matters
for tag, dt, rank in (("lora_bf16_r16", torch.bfloat16, 16), ("lora_bf16_r4", torch.bfloat16, 4), ("lora_f32_r8", torch.float32, 8)):
    base = torch.nn.Linear(64, 48, bias=True).to(dt)
    lay = LoRALinear(LoRAConfig(rank=rank, alpha=2.0, dtype=str(dt).replace("torch.", "")), base)
    torch.nn.init.normal_(lay.lora_up.weight, std=0.05)
    x = torch.randn(2, 5, 64).to(dt)
    G[tag] = {"state": {k: v.clone() for k, v in lay.state_dict().items()}, "x": x, "y": lay(x).detach(), "alpha": 2.0}

# --- norms
x = torch.randn(3, 7, 128).to(torch.bfloat16)
rms = FP32RMSNorm(128, eps=1e-6)
rms.weight.data = torch.randn(128) * 0.2 + 1.0
rms = rms.to(torch.bfloat16)
ln = FP32LayerNorm(128, elementwise_affine=False, eps=1e-5)
G["rmsnorm"] = {"x": x, "w": rms.weight.data.clone(), "y": rms(x).detach()}
G["layernorm"] = {"x": x, "y": ln(x).detach()}
ada = SingleAdaLayerNormZero(128, 128, 32)
for p in ada.parameters():
    torch.nn.init.normal_(p, std=0.3)
ada = ada.to(torch.bfloat16)
te = torch.randn(3, 32).to(torch.bfloat16)
out = ada(x, te)
G["adaln"] = {"x": x, "scale": out.scale.detach(), "shift": out.shift.detach(), "gate": out.gate.detach(),
              "y": out.hidden_states.detach(), "gated": (x + out.hidden_states * out.gate.unsqueeze(1)).detach()}

# --- RoPE tables + application
cfgd = dict(patch_size=16, in_channels=3, out_channels=3, hidden_size=128, depth=2, num_heads=2, mlp_ratio=4.0,
            num_time_tokens=4, rope_theta=256.0, rope_axes_dims=[16, 24, 24], rope_axes_lens=[256, 128, 128],
            context_dim=32, context_start_block=1, do_context_fuse=False, timestep_scale=1.0)
cfg = DenoiserConfig(**cfgd)
emb = RopeEmbedder(rope_theta=256.0, axes_dims=[16, 24, 24], axes_lens=[256, 128, 128], zero_centered=[False, True, True])
H_IMG, W_IMG, CTX = 32, 48, 8
parts = [emb(emb.prepare_image_position_ids(H_IMG, W_IMG, 16, 3).unsqueeze(0)),
         emb(emb.prepare_context_position_ids(6, 2).unsqueeze(0)), emb(emb.prepare_context_position_ids(4, 1).unsqueeze(0)),
         emb(emb.prepare_context_position_ids(CTX, 0).unsqueeze(0))]
freqs = torch.cat(parts, dim=1)                      # [1, L, 32]
L = freqs.shape[1]
q = torch.randn(2, 2, L, 64).to(torch.bfloat16)
G["rope"] = {"freqs_cis": freqs[0].clone(), "x": q, "y": apply_rope(q, freqs.repeat(2, 1, 1)), "height": H_IMG,
             "width": W_IMG, "ctx": CTX, "cfg": cfgd}

# --- attention with the JiT key-padding mask (CPU SDPA, bf16)
k = torch.randn(2, 2, L, 64).to(torch.bfloat16)
v = torch.randn(2, 2, L, 64).to(torch.bfloat16)
km = torch.ones(2, L)
km[0, L - 3:] = 0
km[1, L - 6:] = 0
mask = km.bool().view(2, 1, 1, L).expand(-1, 2, L, -1)
G["attention"] = {"q": q, "k": k, "v": v, "key_mask": km, "y": scaled_dot_product_attention(q, k, v, mask=mask)}

# --- one JiTBlock
blk = JiTBlock(hidden_dim=128, num_heads=2)
for n, p in blk.named_parameters():
    torch.nn.init.normal_(p, std=0.05) if p.dim() > 1 else torch.nn.init.normal_(p, mean=(1.0 if "norm" in n else 0.0), std=0.1)
xb = torch.randn(2, L, 128)
for tag, dt in (("block_f32", torch.float32), ("block_bf16", torch.bfloat16)):
    b2 = blk.to(dt)
    G[tag] = {"state": {k_: v_.clone() for k_, v_ in b2.state_dict().items()}, "x": xb.to(dt), "key_mask": km,
              "y": b2(xb.to(dt), freqs.repeat(2, 1, 1), km).detach()}

# --- whole denoiser, plain and LoRA-wrapped (fp32), with gradients of the rectified-flow loss
torch.manual_seed(99)
den = Denoiser(cfg)
den.initialize_weights()
img = torch.randn(2, 3, H_IMG, W_IMG)
t = torch.sigmoid(torch.randn(2) * 0.8 - 0.8)
ctxt = torch.randn(2, CTX, 32) * 0.5
cmask = torch.ones(2, CTX)
cmask[0, 5:] = 0
cmask[1, 3:] = 0
sizes = torch.tensor([[H_IMG, W_IMG]]).repeat(2, 1)
zeros = torch.zeros_like(sizes)
inputs = dict(image=img, timestep=t, context=ctxt, original_size=sizes, target_size=sizes, crop_coords=zeros, context_mask=cmask)
G["denoiser_f32"] = {"cfg": cfgd, "state": {k_: v_.clone() for k_, v_ in den.state_dict().items()}, "inputs": inputs,
                     "y": den(**inputs).detach()}
peft = PeftTargetConfig(include_keys=[RegexMatch(regex=r"blocks\.\d+\.(attn|mlp)\.")], config=LoRAConfig(rank=16, alpha=4.0, dtype="float32"))
den.requires_grad_(False)
peft.replace_to_peft_layer(den)
for n, p in den.named_parameters():
    if "lora_up" in n:
        torch.nn.init.normal_(p, std=0.05)
        p.requires_grad_(True)
    if "lora_down" in n:
        p.requires_grad_(True)
clean = torch.randn(2, 3, H_IMG, W_IMG)
noise = torch.randn(2, 3, H_IMG, W_IMG)
tt = t.view(-1, 1, 1, 1)
noisy = tt * clean + (1 - tt) * noise
inputs_l = dict(inputs, image=noisy)
pred = den(**inputs_l)
denom = (1 - tt).clamp_min(0.05)
loss = torch.nn.functional.mse_loss((pred - noisy) / denom, (clean - noisy) / denom)
loss.backward()
# only the adapter tensors are stored; the base weights are those of "denoiser_f32" (`X.weight` -> `X.linear.weight`)
G["denoiser_lora_f32"] = {"cfg": cfgd, "alpha": 4.0, "lora_state": {k_: v_.clone() for k_, v_ in den.state_dict().items() if "lora_" in k_},
                          "inputs": inputs_l, "clean": clean, "y": pred.detach(), "loss": loss.detach(),
                          "grads": {n: p.grad.clone() for n, p in den.named_parameters() if p.grad is not None}}

# --- patchify / unpatchify and the JiT (ph, pw, c) order
im = torch.randn(2, 3, 32, 48).to(torch.bfloat16)
po = patchify(im, 16)
G["patchify"] = {"image": im, "patches": po.patches.clone(), "back": unpatchify(po.patches, po.latent_height, po.latent_width, 16, 3).image.clone()}
pt = torch.randn(2, 6, 768).to(torch.bfloat16)
G["jit_unpatchify"] = {"patches": pt, "image": den._unpatchify(pt, 32, 48).clone()}

# --- sinusoidal timestep embedding
ts = torch.tensor([0.0, 0.25, 1.0, 37.0, 999.0])
G["timestep_embedding"] = {"t": ts, "y": get_timestep_embedding(ts, 256, flip_sin_to_cos=True, downscale_freq_shift=0)}

out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.pt")
torch.save(G, out_path)
print("wrote", out_path, os.path.getsize(out_path) // 1024, "KiB;", ", ".join(G.keys()))
