"""Evaluation metrics of the reference's evaluate_vae.py on the device.

Mirrors src/pti_ldm_vae/utils/eval_metrics.py (compute_psnr, compute_ssim: same names, arguments and per-sample
return values) and adds the fused call evaluate_vae.py:87-98 wants: one kernel pass yields per-sample MSE, MAE, PSNR
and SSIM of the clamped reconstruction / image pair.  CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["compute_psnr", "compute_ssim", "compute_eval_metrics", "ssim_window"]


def ssim_window(device, window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """The reference's normalised 1-D Gaussian (eval_metrics.py:37-41), computed the same way in fp32."""
    coords = torch.arange(window_size, device=device) - window_size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma * sigma))
    return (g / g.sum()).float()


def _check(pred: torch.Tensor, target: torch.Tensor) -> None:
    if pred.shape != target.shape or pred.dim() != 4:
        raise ValueError("pred and target must be [B, C, H, W] tensors of the same shape")


def compute_psnr(pred: torch.Tensor, target: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """PSNR per sample (eval_metrics.py:6-19): 10*log10(data_range^2 / max(mse, 1e-12))."""
    _check(pred, target)
    m = ops.eval_metrics(pred.float(), target.float(), ssim_window(pred.device), clamp=None, data_range=data_range)
    return m[:, 2]


def compute_ssim(pred: torch.Tensor, target: torch.Tensor, data_range: float = 1.0, k1: float = 0.01,
                 k2: float = 0.03) -> torch.Tensor:
    """SSIM per sample with the 11x11 Gaussian window, sigma 1.5, zero padding (eval_metrics.py:22-63)."""
    _check(pred, target)
    m = ops.eval_metrics(pred.float(), target.float(), ssim_window(pred.device), clamp=None, data_range=data_range,
                         k1=k1, k2=k2)
    return m[:, 3]


def compute_eval_metrics(reconstruction: torch.Tensor, images: torch.Tensor, data_range: float = 1.0) -> dict:
    """evaluate_vae.py:87-98 in one pass: clamp both to [0, 1], then per-sample psnr, ssim, mse, mae."""
    _check(reconstruction, images)
    m = ops.eval_metrics(reconstruction.float(), images.float(), ssim_window(images.device), clamp=(0.0, 1.0),
                         data_range=data_range)
    return {"mse": m[:, 0], "mae": m[:, 1], "psnr": m[:, 2], "ssim": m[:, 3]}
