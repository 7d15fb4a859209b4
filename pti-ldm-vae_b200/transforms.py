"""Device-side input normalisation.

LocalNormalizeByMask mirrors src/pti_ldm_vae/data/transforms.py:8-32 -- per image, z-score with the mean and
(population) standard deviation of the NON-ZERO pixels, std <= 1e-5 replaced by 1, zero pixels left at exactly 0 -- but
takes a whole batch that already sits on the GPU and returns a CUDA tensor (the reference works on one numpy image on
the host inside the DataLoader workers).  CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["LocalNormalizeByMask", "ApplyLocalNormd"]


class LocalNormalizeByMask:
    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        """img: CUDA tensor [B, ...] (each leading index is one image) or a single [H, W] image."""
        if not isinstance(img, torch.Tensor):
            raise TypeError("the device transform takes torch tensors (use the reference transform for numpy input)")
        single = img.dim() <= 2
        x = img.float()
        x = x.unsqueeze(0) if single else x
        out = ops.local_normalize(x)
        return out[0] if single else out


class ApplyLocalNormd:
    """Dictionary version (transforms.py:35-62): applies LocalNormalizeByMask to the listed keys."""

    def __init__(self, keys: list) -> None:
        self.keys = keys
        self.norm = LocalNormalizeByMask()

    def __call__(self, data: dict) -> dict:
        for k in self.keys:
            data[k] = self.norm(data[k])
        return data
