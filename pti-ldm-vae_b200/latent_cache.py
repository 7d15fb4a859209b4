"""Batched writer / reader of the reference's per-image latent cache.

The reference's ``LatentCache.get_or_encode_batch`` (/root/reference/src/pti_ldm_vae/analysis/latent_cache.py:139-230)
walks the image list and calls ``encoder_fn(image_path)`` for every miss -- one preprocessing + one batch-1
``encode_deterministic`` + one device->host read per image (analysis/common.py:73-84, :155-156).  This class keeps the
reference's ON-DISK FORMAT bit for bit (``cache_root/{md5(abs weights path _ patch size)[:8]}/{md5(abs image path _
mtime)[:12]}.npz`` with arrays ``latent`` and ``patient_id``, plus ``_metadata.json``), so caches are interchangeable
with the reference's class in both directions, but encodes all misses of a call in batches: ``load_fn`` runs per image on
the host (the reference's MONAI transform), the stack goes to the GPU once per batch and ``encode_batch_fn`` -- by default
``vae.encode_deterministic`` of this repo's model -- runs at the batch size where the kernels are efficient.
"""
from __future__ import annotations

import hashlib
import json
import os
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

__all__ = ["LatentCacheWriter", "patient_id_from_filename", "encoder_from_vae"]


def patient_id_from_filename(filename: str) -> str:
    """``ID_HA_YEAR_MONTH_PATIENT.tif`` -> ``PATIENT`` (analysis/latent_space.py:21-37)."""
    stem = filename.rsplit(".", 1)[0] if "." in filename else filename
    parts = stem.split("_")
    return parts[-1] if parts else stem


def encoder_from_vae(vae, device=None) -> Callable:
    """``encode_batch_fn`` for a VAEModel of this repo: [B,1,H,W] host tensor -> np.ndarray [B, D] (flattened z_mu)."""
    import torch

    def encode(batch):
        dev = device if device is not None else next(vae.parameters()).device
        with torch.no_grad():
            z = vae.encode_deterministic(batch.to(dev, non_blocking=True))
        return z.flatten(start_dim=1).cpu().numpy()
    return encode


class LatentCacheWriter:
    def __init__(self, cache_root: Path | str = Path("cache/latents")) -> None:
        self.cache_root = Path(cache_root)
        self.cache_root.mkdir(parents=True, exist_ok=True)

    # -- keys and paths: latent_cache.py:41-100 -------------------------------------------------------------------
    @staticmethod
    def model_signature(vae_weights: str, patch_size: tuple[int, int]) -> str:
        return hashlib.md5(f"{Path(vae_weights).resolve()}_{patch_size}".encode()).hexdigest()[:8]

    @staticmethod
    def image_cache_key(image_path: str) -> str:
        p = Path(image_path).resolve()
        mtime = p.stat().st_mtime if p.exists() else 0
        return hashlib.md5(f"{p}_{mtime}".encode()).hexdigest()[:12]

    def cache_file_path(self, image_path: str, model_signature: str) -> Path:
        d = self.cache_root / model_signature
        d.mkdir(parents=True, exist_ok=True)
        return d / f"{self.image_cache_key(image_path)}.npz"

    def _metadata_path(self, model_signature: str) -> Path:
        return self.cache_root / model_signature / "_metadata.json"

    def load_metadata(self, model_signature: str) -> dict:
        p = self._metadata_path(model_signature)
        if p.exists():
            with open(p) as f:
                return json.load(f)
        return {"images": {}}

    def save_metadata(self, model_signature: str, metadata: dict) -> None:
        with open(self._metadata_path(model_signature), "w") as f:
            json.dump(metadata, f, indent=2)

    # -- the batched pass -------------------------------------------------------------------------------------------
    def get_or_encode_batch(self, image_paths: Sequence[str], load_fn: Callable, encode_batch_fn: Callable, vae_weights: str,
                            patch_size: tuple[int, int], group_name: str = "", batch_size: int = 64,
                            id_fn: Callable = patient_id_from_filename, verbose: bool = False):
        """Returns ``(latents [n, D], ids, paths)`` in input order, exactly what the reference's method returns.

        load_fn(path) -> tensor [C,H,W] (host);  encode_batch_fn(tensor [B,C,H,W]) -> np.ndarray [B, D]."""
        import torch
        sig = self.model_signature(vae_weights, patch_size)
        metadata = self.load_metadata(sig)
        n = len(image_paths)
        latents: list = [None] * n
        ids: list = [None] * n
        todo = []
        for i, path in enumerate(image_paths):
            f = self.cache_file_path(path, sig)
            meta = metadata["images"].get(str(Path(path).resolve()), {})
            if f.exists() and meta.get("cache_key") == self.image_cache_key(path):
                try:
                    data = np.load(f)
                    latents[i], ids[i] = data["latent"], str(data["patient_id"])
                    continue
                except Exception:  # noqa: BLE001  corrupted entry: re-encode (latent_cache.py:189-194)
                    pass
            todo.append(i)
        for lo in range(0, len(todo), batch_size):
            idx = todo[lo:lo + batch_size]
            batch = torch.stack([load_fn(image_paths[i]) for i in idx])
            if batch.device.type == "cpu" and torch.cuda.is_available():
                batch = batch.pin_memory()
            z = np.asarray(encode_batch_fn(batch))
            if z.shape[0] != len(idx):
                raise ValueError(f"encode_batch_fn returned {z.shape[0]} latents for a batch of {len(idx)}")
            for row, i in zip(z, idx):
                path = image_paths[i]
                pid = id_fn(os.path.basename(path))
                np.savez(self.cache_file_path(path, sig), latent=row, patient_id=pid)
                metadata["images"][str(Path(path).resolve())] = {"cache_key": self.image_cache_key(path), "patient_id": pid}
                latents[i], ids[i] = row, pid
        if todo:
            metadata["model"] = str(Path(vae_weights).name)
            metadata["patch_size"] = list(patch_size)
            self.save_metadata(sig, metadata)
        if verbose:
            print(f"{group_name}: {n - len(todo)} from cache, {len(todo)} newly encoded (cache sig {sig})")
        return np.array(latents), ids, list(image_paths)
