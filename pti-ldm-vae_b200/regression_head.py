"""Kernel-backed MLP head for the latent regression path.

What the reference does there (/root/reference/src/pti_ldm_vae/models/regression_head.py:119-138) is
``frozen VAE.encode_deterministic -> flatten -> Linear/activation/Dropout stack``.  The composition class of the reference
(``VAELatentRegressor``) needs nothing from MONAI and works unchanged on top of this package's ``VAEModel`` (it only calls
``vae.parameters()``, ``vae.eval()``, ``vae.encode_deterministic`` and reads ``vae.autoencoder.in_channels``), so it is NOT
mirrored here.  This module supplies only the piece that runs on the device:

* ``LatentRegressor`` -- holds the ``mlp.<i>.weight / bias`` parameters in the reference's ``nn.Sequential`` positions (so
  ``head_best.pth`` checkpoints load) and evaluates every ``Linear (+ activation)`` pair as ONE ``ptivae_linear_act``
  launch under ``torch.no_grad()`` (eval-mode dropout is the identity); with grad enabled (training the head over the
  frozen encoder) the same ``nn.Sequential`` runs as ordinary autograd ops.
* ``regress_from_images(vae, head, images)`` -- the three-step composition as a function, for callers without the
  reference package on their path.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import nn

from . import ops

ACTIVATIONS = {"relu": nn.ReLU, "gelu": nn.GELU, "leaky_relu": nn.LeakyReLU, "elu": nn.ELU}   # the kernel's epilogues


class LatentRegressor(nn.Module):
    def __init__(self, in_features: int, hidden_dims: Sequence[int], output_dim: int, dropout: float = 0.0,
                 activation: str = "relu") -> None:
        super().__init__()
        if in_features <= 0 or output_dim <= 0:
            raise ValueError(f"in_features ({in_features}) and output_dim ({output_dim}) must be positive")
        if activation not in ACTIVATIONS:
            raise ValueError(f"activation {activation!r} has no fused epilogue; choose one of {sorted(ACTIVATIONS)}")
        widths = [int(in_features), *[int(h) for h in hidden_dims], int(output_dim)]
        # Sequential positions are the checkpoint contract: Linear, activation[, Dropout] per hidden layer, final Linear
        stack: list[nn.Module] = []
        for fan_in, fan_out in zip(widths[:-2], widths[1:-1]):
            stack += [nn.Linear(fan_in, fan_out), ACTIVATIONS[activation]()]
            if dropout > 0:
                stack.append(nn.Dropout(dropout))
        stack.append(nn.Linear(widths[-2], widths[-1]))
        self.mlp = nn.Sequential(*stack)
        self.activation = activation

    def forward(self, latent_flat: torch.Tensor) -> torch.Tensor:
        if not latent_flat.is_cuda:
            raise RuntimeError("LatentRegressor runs on CUDA only (no CPU fallback; the CPU restatement is in oracle/)")
        if torch.is_grad_enabled() and (self.training or latent_flat.requires_grad):
            # TRAINING the head (reg_scripts: frozen encoder -> trainable 1.1 M-parameter MLP with dropout): the 2 MFLOP
            # stack and its backward are ordinary autograd ops on the GPU -- the hot path of that loop is the frozen
            # encoder (17.5 GFLOP per image), which runs on the kernels under no_grad.  Same parameters, same
            # nn.Sequential, so checkpoints and optimizers are shared with the kernel-backed inference path below.
            return self.mlp(latent_flat)
        linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
        y = latent_flat
        with torch.cuda.device(latent_flat.device):
            for k, lin in enumerate(linears):
                last = k == len(linears) - 1
                y = ops.linear_act(y, lin.weight, lin.bias, None if last else self.activation)
        return y


def regress_from_images(vae, head: LatentRegressor, images: torch.Tensor) -> torch.Tensor:
    """mu = vae.encode_deterministic(images) -> flatten -> head (regression_head.py:119-138 as a function)."""
    with torch.no_grad():
        mu = vae.encode_deterministic(images)
    flat = torch.flatten(mu, start_dim=1)
    first = next(m for m in head.mlp if isinstance(m, nn.Linear))
    if first.in_features != flat.shape[1]:
        raise ValueError(f"the head expects {first.in_features} latent features, the encoder produced {flat.shape[1]} "
                         f"(latent {tuple(mu.shape[1:])})")
    return head(flat)
