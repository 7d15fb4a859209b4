"""Checkpoint contract of the reference (src/pti_ldm_vae/utils/vae_loader.py:27-43): a file holding
either a bare state_dict or a dict with ``autoencoder_state_dict``; keys carry no ``autoencoder.``
prefix."""
from __future__ import annotations

import torch

from .vae_model import VAEModel


def load_vae_model(autoencoder_def: dict, checkpoint_path, device) -> VAEModel:
    model = VAEModel.from_config(autoencoder_def).to(device)
    ckpt = torch.load(checkpoint_path, map_location=device, weights_only=True)
    if isinstance(ckpt, dict) and "autoencoder_state_dict" in ckpt:
        ckpt = ckpt["autoencoder_state_dict"]
    model.load_state_dict(ckpt)
    model.eval()
    return model
