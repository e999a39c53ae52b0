"""TREAD token routing: a random subset of the patch tokens skips the middle blocks and is re-inserted afterwards.

Mirror of /root/reference/train/jit/class_to_image_tread.py:73-118 (`keep_and_route_tokens`) and of the re-insertion at
the end of routing.  The gathers / scatters of [B, L, D] token rows are one 16-byte-vectorised kernel each
(`vpt_token_gather`), with autograd; the RoPE table rows and the mask columns of the kept tokens are index selects of
small tensors."""
from __future__ import annotations

import torch

from ... import ops


def keep_and_route_tokens(patch_tokens: torch.Tensor, cos_sin: torch.Tensor, mask: torch.Tensor, route_rate: float,
                          perm: torch.Tensor | None = None):
    """Returns (keep_tokens, route_tokens, keep_cos_sin, route_cos_sin, keep_mask, route_mask, inverse_perm).
    `num_keep = int(L * route_rate)` as in the reference; `cos_sin` is the per-token table [L, hd/2, 2] (batch-invariant
    here), `mask` [B, L].  `perm` may be given for reproducibility (default: torch.randperm on the tokens' device)."""
    _, L, _ = patch_tokens.shape
    num_keep = int(L * route_rate)
    if perm is None:
        perm = torch.randperm(L, device=patch_tokens.device)
    keep_idx, route_idx = perm[:num_keep].contiguous(), perm[num_keep:].contiguous()
    inverse_perm = torch.argsort(perm)
    return (ops.token_gather(patch_tokens, keep_idx), ops.token_gather(patch_tokens, route_idx), cos_sin[keep_idx], cos_sin[route_idx],
            mask[:, keep_idx], mask[:, route_idx], inverse_perm)


def merge_routed_tokens(keep_tokens: torch.Tensor, route_tokens: torch.Tensor, inverse_perm: torch.Tensor) -> torch.Tensor:
    """cat([keep, route], dim=1)[:, inverse_perm] without the concatenated copy: two scatters into the full buffer."""
    n_keep = keep_tokens.shape[1]
    perm = torch.argsort(inverse_perm)                       # position in the full sequence of every kept / routed token
    return ops.token_merge(keep_tokens, route_tokens, perm[:n_keep].contiguous(), perm[n_keep:].contiguous())
