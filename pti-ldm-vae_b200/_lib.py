"""ctypes binding of libptivae.so (the C ABI declared in include/ptivae.h).

There is NO CPU fallback: if the shared library is missing or fails to load, every op raises.
"""
from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess

_PKG_DIR = pathlib.Path(__file__).resolve().parent
_CSRC = _PKG_DIR / "csrc"
LIB_PATH = _PKG_DIR / "libptivae.so"

_c_void_p, _c_int, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_c_ll, _c_ull = ctypes.c_longlong, ctypes.c_ulonglong

# symbol -> argtypes ; must list every function include/ptivae.h declares (tests check this)
SIGNATURES = {
    "ptivae_abi_version": [],
    "ptivae_conv_umma": [_c_void_p] * 7 + [_c_int] * 10 + [_c_void_p],
    "ptivae_conv_parts": [_c_int] * 3,
    "ptivae_conv3x3_fused": [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p,
                             _c_int, _c_void_p] + [_c_int] * 8 + [_c_void_p],
    "ptivae_set_chained_launch": [_c_int],
    "ptivae_conv3x3_fused_parts": [_c_int] * 2,
    "ptivae_conv3x3_fused_query": [_c_int] * 6,
    "ptivae_conv3x3_fused_sc": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p,
                                _c_int, _c_void_p] + [_c_int] * 7 + [_c_void_p],
    "ptivae_up2x_conv3x3": [_c_void_p] * 6 + [_c_int] * 6 + [_c_void_p],
    "ptivae_up2x_conv3x3_parts": [_c_int] * 2,
    "ptivae_debug_set_trace": [_c_void_p],
    "ptivae_eval_metrics": [_c_void_p] * 3 + [_c_int] + [_c_void_p] * 2 + [_c_int] * 5 + [_c_float] * 5 + [_c_void_p],
    "ptivae_eval_metrics_workspace": [_c_int] * 4,
    "ptivae_local_normalize": [_c_void_p] * 4 + [_c_int] * 2 + [_c_void_p],
    "ptivae_local_normalize_workspace": [_c_int],
    "ptivae_resize_area": [_c_void_p, _c_int, _c_void_p] + [_c_int] * 5 + [_c_void_p],
    "ptivae_pack_conv_weight": [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p],
    "ptivae_gn_stats": [_c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p],
    "ptivae_gn_stats_parts": [_c_int] * 3,
    "ptivae_gn_finalize": [_c_void_p] * 5 + [_c_int] * 5 + [_c_float, _c_void_p],
    "ptivae_gn_apply": [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p],
    "ptivae_gn_finalize_checked": [_c_void_p] * 5 + [_c_int] * 5 + [_c_float, _c_void_p, _c_void_p],
    "ptivae_range_check": [_c_void_p, _c_ll, _c_void_p, _c_void_p],
    "ptivae_conv3x3_small_cin": [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p],
    "ptivae_conv3x3_small_cin_parts": [_c_int] * 3,
    "ptivae_conv3x3_small_cout": [_c_void_p] * 5 + [_c_int] * 6 + [_c_void_p],
    "ptivae_conv1x1_small": [_c_void_p] * 4 + [_c_int] * 5 + [_c_void_p],
    "ptivae_attention_fwd": [_c_void_p] * 5 + [_c_int] * 5 + [_c_void_p],
    "ptivae_latent_sample": [_c_void_p] * 6 + [_c_ll, _c_ull, _c_ull, _c_void_p],
    "ptivae_rng_advance": [_c_void_p, _c_void_p],
    "ptivae_kl_loss": [_c_void_p] * 4 + [_c_int] * 3 + [_c_void_p],
    "ptivae_l1l2": [_c_void_p] * 4 + [_c_ll, _c_void_p],
    "ptivae_spatial_mean": [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "ptivae_ar_vae_loss": [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p] * 4,
    "ptivae_ar_vae_loss_bwd": [_c_void_p] * 5 + [_c_int] * 4 + [_c_void_p] * 5,
    "ptivae_spatial_mean_bwd": [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "ptivae_linear_act": [_c_void_p] * 4 + [_c_int] * 4 + [_c_void_p],
    # backward pass
    "ptivae_wgrad": [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p],
    "ptivae_wgrad_workspace": [_c_int] * 6,
    "ptivae_debug_set_wgrad_status": [_c_void_p],
    "ptivae_bgemm": [_c_void_p] * 3 + [_c_int] * 4 + [_c_ll, _c_ll, _c_int, _c_int, _c_ll, _c_ll, _c_int, _c_int, _c_ll, _c_ll,
                                                      _c_int, _c_int, _c_float, _c_void_p, _c_void_p, _c_ll, _c_ll, _c_int,
                                                      _c_void_p],
    "ptivae_rowdot": [_c_void_p] * 3 + [_c_ll, _c_int, _c_ll, _c_ll, _c_int, _c_int, _c_void_p],
    "ptivae_gn_bwd": [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int] +
                     [_c_void_p] * 8 + [_c_int] * 5 + [_c_void_p],
    "ptivae_gn_bwd_workspace": [_c_int] * 3,
    "ptivae_colsum": [_c_void_p] * 3 + [_c_ll, _c_int, _c_int, _c_void_p],
    "ptivae_colsum_blocks": [_c_ll],
    "ptivae_thin_wgrad": [_c_void_p] * 6 + [_c_int] * 7 + [_c_void_p],
    "ptivae_thin_wgrad_workspace": [_c_int] * 5,
    "ptivae_latent_bwd": [_c_void_p] * 15 + [_c_int] * 3 + [_c_void_p],
    "ptivae_outer_reduce": [_c_void_p] * 4 + [_c_int] * 4 + [_c_void_p],
    "ptivae_l1l2_bwd": [_c_void_p] * 4 + [_c_ll, _c_void_p],
    "ptivae_kl_bwd": [_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p],
    "ptivae_pack_many": [_c_void_p, _c_int, _c_ll, _c_void_p],
    "ptivae_pack_desc_bytes": [],
    "ptivae_cast16": [_c_void_p, _c_void_p, _c_ll, _c_int, _c_int, _c_void_p],
    "ptivae_adam": [_c_void_p] * 4 + [_c_ll] + [_c_float] * 5 + [_c_void_p, _c_int, _c_void_p],
}
_RESTYPE_LL = {"ptivae_wgrad_workspace", "ptivae_thin_wgrad_workspace", "ptivae_gn_bwd_workspace"}

_lib = None


def build(verbose: bool = False) -> pathlib.Path:
    """Compile csrc/*.cu for sm_100a into libptivae.so (in-tree; nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", str(_CSRC), "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libptivae.so failed (see output above)")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C pti-ldm-vae_b200/csrc`.")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export it
            fn.argtypes = argtypes
            fn.restype = ctypes.c_longlong if name in _RESTYPE_LL else ctypes.c_int
        _lib = handle
    return _lib


class PtivaeError(RuntimeError):
    pass


_ERRS = {-1: "bad argument", -2: "unsupported shape", -3: "driver entry point / tensor-map encode failed"}


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = _ERRS.get(rc, f"cudaError {rc}" if rc > 0 else f"error {rc}")
        raise PtivaeError(f"{what}: {msg}")
