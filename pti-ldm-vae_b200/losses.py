"""Loss entry points with the reference's names and semantics
(/root/reference/src/pti_ldm_vae/models/losses.py:4-66; L1/MSE: vae_scripts/train_vae.py:289-296).
The reductions run as deterministic two-stage CUDA kernels (csrc/latent_loss.cu)."""
from __future__ import annotations

import torch

from . import ops


def compute_kl_loss(z_mu: torch.Tensor, z_logvar: torch.Tensor, *, input_is_logvar: bool = True) -> torch.Tensor:
    """KL of a diagonal Gaussian, batch mean.  As the reference calls it (train_vae.py:394) the second
    argument is sigma interpreted as log-variance; that quirk is preserved bit-for-formula."""
    return ops.kl_loss(z_mu, z_logvar, input_is_logvar)


def l1_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ops.l1l2(a, b)[0]


def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return ops.l1l2(a, b)[1]


def l1_and_mse(a: torch.Tensor, b: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    r = ops.l1l2(a, b)
    return r[0], r[1]


def compute_total_loss(recons_loss, kl_loss, perceptual_loss, adv_gen_loss, ar_loss, *, kl_weight: float,
                       perceptual_weight: float, adv_weight: float, ar_gamma: float, ar_vae_enabled: bool):
    """losses.py:33-66 -- scalar arithmetic, stays in PyTorch."""
    total = recons_loss + kl_weight * kl_loss + perceptual_weight * perceptual_loss + adv_weight * adv_gen_loss
    if ar_vae_enabled:
        total = total + ar_gamma * ar_loss
    return total
