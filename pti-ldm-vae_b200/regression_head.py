"""Latent regression head with the reference's classes and semantics
(/root/reference/src/pti_ldm_vae/models/regression_head.py:30-169): ``LatentRegressor`` (MLP over flattened
latents) and ``VAELatentRegressor`` (frozen VAE ``encode_deterministic`` -> flatten -> MLP).  Module tree and
state_dict keys (``mlp.<i>.weight``...) match the reference, so ``head_best.pth`` checkpoints load.  On CUDA the
forward runs the sm_100a kernels (encoder path of AutoencoderKL + ``ptivae_linear_act``); inference only.
"""
from __future__ import annotations

import warnings
from collections.abc import Iterable, Sequence

import torch
from torch import nn

from . import ops
from .vae_model import VAEModel

_ACT_MODULES = {"relu": nn.ReLU, "gelu": nn.GELU, "leaky_relu": nn.LeakyReLU, "elu": nn.ELU}


def _activation_from_name(name: str) -> nn.Module:
    if name not in _ACT_MODULES:
        raise ValueError(f"Unsupported activation: {name}. Choose from {', '.join(_ACT_MODULES)}.")
    return _ACT_MODULES[name]()


class LatentRegressor(nn.Module):
    def __init__(self, in_features: int, hidden_dims: Sequence[int], output_dim: int, dropout: float = 0.0,
                 activation: str = "relu") -> None:
        super().__init__()
        if in_features <= 0:
            raise ValueError("in_features must be positive.")
        if output_dim <= 0:
            raise ValueError("output_dim must be positive.")
        layers: list[nn.Module] = []
        dims = [in_features, *hidden_dims, output_dim]
        act = _activation_from_name(activation)
        for idx in range(len(dims) - 2):
            layers.append(nn.Linear(dims[idx], dims[idx + 1]))
            layers.append(act.__class__())
            if dropout > 0:
                layers.append(nn.Dropout(p=dropout))
        layers.append(nn.Linear(dims[-2], dims[-1]))
        self.mlp = nn.Sequential(*layers)
        self._act_name = activation

    def forward(self, latent_flat: torch.Tensor) -> torch.Tensor:
        if not latent_flat.is_cuda:
            raise RuntimeError("LatentRegressor runs on CUDA only (no CPU fallback; the CPU restatement is in oracle/)")
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("backward kernels are not built yet: call under torch.no_grad() / .eval()")
        x = latent_flat
        mods = list(self.mlp)
        i = 0
        while i < len(mods):
            lin = mods[i]
            assert isinstance(lin, nn.Linear)
            has_act = i + 1 < len(mods) and not isinstance(mods[i + 1], (nn.Linear, nn.Dropout))
            x = ops.linear_act(x, lin.weight, lin.bias, self._act_name if has_act else None)
            i += 1 + int(has_act)
            if i < len(mods) and isinstance(mods[i], nn.Dropout):
                i += 1          # eval-mode dropout is the identity
        return x


class VAELatentRegressor(nn.Module):
    def __init__(self, vae: VAEModel, regressor: LatentRegressor, *, latent_dim: int,
                 flatten_warning_threshold: int = 131072) -> None:
        super().__init__()
        self.vae = vae
        self.regressor = regressor
        self.latent_dim = latent_dim
        first_linear = next((layer for layer in self.regressor.mlp if isinstance(layer, nn.Linear)), None)
        if first_linear is None or first_linear.in_features != latent_dim:
            raise ValueError(f"Regression head expects in_features={latent_dim}, "
                             f"got {first_linear.in_features if first_linear else 'unknown'}.")
        for param in self.vae.parameters():
            param.requires_grad = False
        self.vae.eval()
        self.flatten_warning_threshold = flatten_warning_threshold

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            latent = self.vae.encode_deterministic(images)
        latent_flat = torch.flatten(latent, start_dim=1)
        if latent_flat.shape[1] > self.flatten_warning_threshold:
            warnings.warn(f"Flattened latent dimension {latent_flat.shape[1]} is large; consider reducing patch size "
                          "or latent channels.", stacklevel=2)
        return self.regressor(latent_flat)

    @staticmethod
    def compute_flat_dim(latent: torch.Tensor) -> int:
        return int(torch.flatten(latent, start_dim=1).shape[1])

    @staticmethod
    def infer_flat_dim_from_patch(vae: VAEModel, patch_size: Iterable[int], device: torch.device, *,
                                  channels: int | None = None) -> int:
        height, width = patch_size
        inferred_channels = channels if channels is not None else getattr(vae.autoencoder, "in_channels", 1)
        with torch.no_grad():
            dummy = torch.zeros(1, inferred_channels, height, width, device=device)
            latent = vae.encode_deterministic(dummy)
        return VAELatentRegressor.compute_flat_dim(latent)
