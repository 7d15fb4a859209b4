"""Device-side input normalisation.

LocalNormalizeByMask mirrors src/pti_ldm_vae/data/transforms.py:8-32 -- per image, z-score with the mean and
(population) standard deviation of the NON-ZERO pixels, std <= 1e-5 replaced by 1, zero pixels left at exactly 0 -- but
takes a whole batch that already sits on the GPU and returns a CUDA tensor (the reference works on one numpy image on
the host inside the DataLoader workers).  CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["LocalNormalizeByMask", "ApplyLocalNormd", "read_tiff", "preprocess_batch"]


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


# ---------------------------------------------------------------------------------------------------------------------
# GPU input pipeline: decode on the host, everything else on the device
# ---------------------------------------------------------------------------------------------------------------------
def read_tiff(path) -> "np.ndarray":
    """Minimal baseline-TIFF decoder for the reference's inputs (single-channel uint8 / uint16 / float32 panoramics, read
    there with ``tifffile.imread`` -- data/transforms.py:64-77 -- or MONAI ``LoadImage``): uncompressed strips, little or
    big endian, one sample per pixel.  Returns the pixels in their stored dtype, [H, W].  Anything else (compression,
    tiles, RGB) raises ``ValueError`` -- decode those with the reference's reader and hand the array to
    ``preprocess_batch``."""
    import struct
    import numpy as np
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] == b"II":
        e = "<"
    elif data[:2] == b"MM":
        e = ">"
    else:
        raise ValueError(f"{path}: not a TIFF file")
    if struct.unpack(e + "H", data[2:4])[0] != 42:
        raise ValueError(f"{path}: BigTIFF / unknown TIFF version")
    off = struct.unpack(e + "I", data[4:8])[0]
    n = struct.unpack(e + "H", data[off:off + 2])[0]
    sizes = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 16: 8}
    fmts = {1: "B", 3: "H", 4: "I", 16: "Q"}
    tags = {}
    for i in range(n):
        ent = data[off + 2 + 12 * i: off + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(e + "HHI", ent[:8])
        if typ not in fmts:
            continue
        nbytes = sizes[typ] * cnt
        raw = ent[8:8 + nbytes] if nbytes <= 4 else data[struct.unpack(e + "I", ent[8:12])[0]:][:nbytes]
        tags[tag] = struct.unpack(e + fmts[typ] * cnt, raw)
    w, h = tags[256][0], tags[257][0]
    bits = tags.get(258, (1,))[0]
    if tags.get(259, (1,))[0] != 1:
        raise ValueError(f"{path}: compressed TIFF (compression tag {tags[259][0]}) is not handled by this reader")
    if tags.get(277, (1,))[0] != 1 or 322 in tags:
        raise ValueError(f"{path}: multi-sample or tiled TIFF is not handled by this reader")
    fmt = tags.get(339, (1,))[0]          # 1 unsigned int, 3 IEEE float
    dt = {(8, 1): "u1", (16, 1): "u2", (32, 3): "f4", (32, 1): "u4"}.get((bits, fmt))
    if dt is None:
        raise ValueError(f"{path}: {bits}-bit sample format {fmt} is not handled by this reader")
    offs, cnts = tags[273], tags.get(279)
    if cnts is None:
        cnts = (h * w * bits // 8,)
    buf = b"".join(data[o:o + c] for o, c in zip(offs, cnts))
    img = np.frombuffer(buf, dtype=np.dtype(dt).newbyteorder(e), count=h * w).reshape(h, w)
    return img.astype(np.dtype(dt))      # native byte order


def preprocess_batch(raw, patch_size, device=None) -> torch.Tensor:
    """The reference's per-image preprocessing (dataloaders.py:263-272: Resize(patch_size) -> LocalNormalizeByMask ->
    float32, channel first) for a whole batch on the GPU.

    raw: list of equally sized [H, W] numpy arrays / tensors, or one [B, H, W] array / tensor, uint8 / uint16 / float32
    (what ``read_tiff`` or the reference's readers return).  Returns a CUDA fp32 tensor [B, 1, h, w]."""
    import numpy as np
    if isinstance(raw, (list, tuple)):
        raw = np.stack([np.asarray(r) for r in raw]) if not isinstance(raw[0], torch.Tensor) else torch.stack(list(raw))
    if isinstance(raw, np.ndarray):
        if raw.dtype == np.uint16:
            raw = raw.view(np.int16)          # same bits; the kernel reads them as unsigned
        elif raw.dtype not in (np.uint8, np.float32):
            raw = raw.astype(np.float32)
        raw = torch.from_numpy(np.ascontiguousarray(raw))
    if raw.dim() != 3:
        raise ValueError(f"expected [B, H, W] raw images, got {tuple(raw.shape)}")
    if device is None:
        device = raw.device if raw.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if not raw.is_cuda:
        raw = raw.pin_memory().to(device, non_blocking=True)
    with torch.cuda.device(raw.device):
        x = ops.resize_area(raw, patch_size)
        return ops.local_normalize(x).unsqueeze(1)
