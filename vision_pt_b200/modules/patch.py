"""patchify / unpatchify -- mirror of /root/reference/src/modules/patch.py:7-174 on the coalesced permute kernel."""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from .. import ops


class PatchifyOutput(NamedTuple):
    patches: torch.Tensor
    latent_height: int
    latent_width: int


class UnpatchifyOutput(NamedTuple):
    image: torch.Tensor


def patchify(image: torch.Tensor, patch_size: int) -> PatchifyOutput:
    if image.dim() == 3:
        image = image.unsqueeze(0)
    elif image.dim() != 4:
        raise ValueError("Input image must be 3D or 4D tensor")
    _, _, height, width = image.shape
    patches = ops.patchify_op(image, patch_size, order=0)
    return PatchifyOutput(patches=patches, latent_height=height // patch_size, latent_width=width // patch_size)


def unpatchify(patches: torch.Tensor, latent_height: int, latent_width: int, patch_size: int,
               out_channels: int) -> UnpatchifyOutput:
    if patches.dim() == 2:
        patches = patches.unsqueeze(0)
    elif patches.dim() != 3:
        raise ValueError("Input patches must be 2D or 3D tensor")
    image = ops.unpatchify_op(patches, out_channels, latent_height * patch_size, latent_width * patch_size, patch_size,
                              order=0)
    return UnpatchifyOutput(image=image)


class ImagePatcher(nn.Module):
    def __init__(self, patch_size: int, out_channels: int):
        super().__init__()
        self.patch_size = patch_size
        self.out_channels = out_channels

    def patchify(self, image: torch.Tensor) -> PatchifyOutput:
        return patchify(image, self.patch_size)

    def unpatchify(self, patches: torch.Tensor, latent_height: int, latent_width: int) -> UnpatchifyOutput:
        return unpatchify(patches, latent_height, latent_width, self.patch_size, self.out_channels)

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        return self.patchify(image).patches
