"""Config loading compatible with the reference's ``config/*.json`` files.

The reference reads them with MONAI ``ConfigParser`` (vae_scripts/train_vae.py:100-124,
src/pti_ldm_vae/utils/vae_loader.py:11-24).  MONAI is not a dependency here, so this module
implements the subset those files use: a string value that is *entirely* ``@id`` (ids nest with
``::`` or ``#``) is replaced by the referenced value; everything else -- including
``"@regularized_attributes.gamma"``, whose ``.`` is not an id separator -- stays a literal string,
and ``_comment`` keys are carried along.
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Any

_ID = re.compile(r"^@([A-Za-z0-9_]+(?:(?:::|#)[A-Za-z0-9_]+)*)$")


def _lookup(root: Any, ref: str):
    node = root
    for part in re.split(r"::|#", ref):
        if isinstance(node, dict) and part in node:
            node = node[part]
        elif isinstance(node, list) and part.isdigit() and int(part) < len(node):
            node = node[int(part)]
        else:
            raise KeyError(ref)
    return node


def resolve_references(cfg: dict) -> dict:
    def walk(node, depth=0):
        if depth > 32:
            raise ValueError("reference cycle in config")
        if isinstance(node, dict):
            return {k: walk(v, depth) for k, v in node.items()}
        if isinstance(node, list):
            return [walk(v, depth) for v in node]
        if isinstance(node, str):
            m = _ID.match(node)
            if m:
                try:
                    return walk(_lookup(cfg, m.group(1)), depth + 1)
                except KeyError:
                    return node  # unresolved ids stay literal (MONAI would raise; the reference never hits this)
        return node

    return walk(cfg)


def load_config(path: str | Path) -> dict:
    with open(path, encoding="utf-8") as f:
        return resolve_references(json.load(f))


def load_vae_config(path: str | Path) -> dict:
    """Mirror of utils/vae_loader.py:11-24: returns the resolved ``autoencoder_def`` section."""
    return load_config(path)["autoencoder_def"]


# The two autoencoder_def variants of the reference's config/ directory (SURVEY.md 8 "Config A/B").
AUTOENCODER_DEF_A = {  # vae_dente_no_adv / vae_dente_2 / vae_both_no_adv / vae_edente_no_adv
    "spatial_dims": 2, "in_channels": 1, "out_channels": 1, "latent_channels": 4,
    "channels": [32, 64, 128, 128], "num_res_blocks": 2, "norm_num_groups": 16, "norm_eps": 1e-6,
    "attention_levels": [False, False, False, False],
    "with_encoder_nonlocal_attn": True, "with_decoder_nonlocal_attn": True,
}
AUTOENCODER_DEF_B = {  # ar_vae_dente
    "spatial_dims": 2, "in_channels": 1, "out_channels": 1, "latent_channels": 10,
    "channels": [64, 128, 256], "num_res_blocks": 2, "norm_num_groups": 32, "norm_eps": 1e-6,
    "attention_levels": [False, False, False],
    "with_encoder_nonlocal_attn": True, "with_decoder_nonlocal_attn": True,
}
