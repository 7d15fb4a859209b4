"""``VAEModel`` with the reference's public surface
(/root/reference/src/pti_ldm_vae/models/autoencoder.py:6-171): same constructor arguments and
defaults, ``from_config``, ``forward -> (reconstruction, z_mu, z_sigma)``, the stochastic /
deterministic encode helpers, and ``state_dict``/``load_state_dict`` that pass through to the inner
autoencoder WITHOUT an ``autoencoder.`` prefix (autoencoder.py:165-171).  Only the inner module
differs: ``AutoencoderKL`` here is the sm_100a implementation, not MONAI's.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .autoencoderkl import AutoencoderKL


class VAEModel(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, latent_channels: int,
                 channels: list[int], num_res_blocks: int = 2, norm_num_groups: int = 32, norm_eps: float = 1e-6,
                 attention_levels: list[bool] | None = None, with_encoder_nonlocal_attn: bool = True,
                 with_decoder_nonlocal_attn: bool = True) -> None:
        super().__init__()
        if attention_levels is None:
            attention_levels = [False] * len(channels)
        self.autoencoder = AutoencoderKL(
            spatial_dims=spatial_dims, in_channels=in_channels, out_channels=out_channels,
            latent_channels=latent_channels, channels=channels, num_res_blocks=num_res_blocks,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps, attention_levels=attention_levels,
            with_encoder_nonlocal_attn=with_encoder_nonlocal_attn,
            with_decoder_nonlocal_attn=with_decoder_nonlocal_attn)

    @classmethod
    def from_config(cls, config: dict) -> "VAEModel":
        # unknown keys (the "_comment" entries of config/*.json) are ignored, as in the reference
        return cls(
            spatial_dims=config["spatial_dims"], in_channels=config["in_channels"],
            out_channels=config["out_channels"], latent_channels=config["latent_channels"],
            channels=config["channels"], num_res_blocks=config.get("num_res_blocks", 2),
            norm_num_groups=config.get("norm_num_groups", 32), norm_eps=config.get("norm_eps", 1e-6),
            attention_levels=config.get("attention_levels"),
            with_encoder_nonlocal_attn=config.get("with_encoder_nonlocal_attn", True),
            with_decoder_nonlocal_attn=config.get("with_decoder_nonlocal_attn", True))

    def forward(self, x: torch.Tensor):
        """-> (reconstruction, z_mu, z_sigma).  The third output is the standard deviation (the
        reference names it z_logvar, train_vae.py:385); sampling happens in eval mode too."""
        return self.autoencoder(x)

    def encode_stage_2_inputs(self, x: torch.Tensor) -> torch.Tensor:
        return self.autoencoder.encode_stage_2_inputs(x)

    def encode_deterministic(self, x: torch.Tensor) -> torch.Tensor:
        z_mu, _ = self.autoencoder.encode(x)
        return z_mu

    def decode_stage_2_outputs(self, z: torch.Tensor) -> torch.Tensor:
        return self.autoencoder.decode_stage_2_outputs(z)

    def reconstruct_deterministic(self, x: torch.Tensor) -> torch.Tensor:
        return self.decode_stage_2_outputs(self.encode_deterministic(x))

    def load_state_dict(self, state_dict: dict, strict: bool = True):
        return self.autoencoder.load_state_dict(state_dict, strict=strict)

    def state_dict(self, *args, **kwargs) -> dict:
        return self.autoencoder.state_dict(*args, **kwargs)
