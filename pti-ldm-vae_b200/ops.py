"""Thin torch-tensor wrappers over the C ABI (include/ptivae.h).

PyTorch is plumbing here: it owns device memory and the current stream; every op below hands raw
device pointers to libptivae.so.  No op has a CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib

BF16 = torch.bfloat16
F16 = torch.float16
_FMT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}   # storage codes of the C ABI


def _fmt(t: torch.Tensor) -> int:
    try:
        return _FMT[t.dtype]
    except KeyError:
        raise _lib.PtivaeError(f"unsupported storage dtype {t.dtype}") from None


def _op16(t: torch.Tensor) -> int:
    """f16 flag of a 16-bit GEMM operand tensor."""
    if t.dtype not in (BF16, F16):
        raise _lib.PtivaeError(f"GEMM operands must be float16 or bfloat16, got {t.dtype}")
    return int(t.dtype == F16)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


# --- instrumentation used by bench.py: launch counting and per-op CUDA-event timing -------------
LAUNCHES = 0          # kernels launched through this module since import
PROFILE = None        # set to a list to record (name, meta, start_event, end_event) per C call


def _call(name: str, meta, kernels: int, fn, *args) -> None:
    global LAUNCHES
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.append((name, meta, e0, e1))
    else:
        rc = fn(*args)
    _lib.check(rc, f"{name}{meta if meta is not None else ''}")
    LAUNCHES += kernels


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PtivaeError("ptivae ops need CUDA tensors (there is no CPU path)")


def pack_conv_weight(w: torch.Tensor, mode: int = 0, dtype: torch.dtype = F16) -> torch.Tensor:
    """fp32 [Cout,Cin,k,k] (or Linear [out,in]) -> 16-bit [T,Cout,Cin]."""
    _need_cuda(w)
    if w.dim() == 2:
        w = w[:, :, None, None]
    w = w.detach().contiguous().float()
    cout, cin, k, _ = w.shape
    t = 16 if mode == 2 else k * k
    out = torch.empty((t, cout, cin), device=w.device, dtype=dtype)
    _call("pack_conv_weight", None, 1, _lib.lib().ptivae_pack_conv_weight, _p(w), _p(out), cout, cin, k, mode,
          _op16(out), _stream())
    return out


def conv_parts(h: int, w: int, mode: int) -> int:
    return _lib.lib().ptivae_conv_parts(h, w, mode)


def conv_umma(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, mode: int, residual=None,
              gn_groups: int = 0, out_f32: bool = False, emit16: bool = False):
    """x bf16 NHWC [N,H,W,Cin] -> NHWC (bf16, or fp32 when out_f32).  mode: 0 3x3 s1 | 1 pad+3x3 s2 |
    2 up2x+3x3 | 3 1x1.  Returns out, or (out, stats_partials [N,P,G,2]) when gn_groups > 0; with
    emit16 (fp32 output only) a 16-bit copy of out is appended to the returned tuple."""
    _need_cuda(x, w_packed, bias)
    f16 = _op16(x)
    if w_packed.dtype != x.dtype:
        raise _lib.PtivaeError("activation and packed-weight operand dtypes differ")
    n, h, w, cin = x.shape
    cout = w_packed.shape[1]
    ho, wo = (h // 2, w // 2) if mode == 1 else ((2 * h, 2 * w) if mode == 2 else (h, w))
    out = torch.empty((n, ho, wo, cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    res_f32 = 0
    if residual is not None:
        if residual.shape != out.shape:
            raise _lib.PtivaeError("residual shape mismatch")
        res_f32 = int(residual.dtype == torch.float32)
        if not res_f32 and residual.dtype != x.dtype:
            raise _lib.PtivaeError("16-bit residual must use the operand dtype")
    part = None
    if gn_groups > 0:
        part = torch.empty((n, conv_parts(h, w, mode), gn_groups, 2), device=x.device, dtype=torch.float32)
    out16 = torch.empty(out.shape, device=x.device, dtype=x.dtype) if (emit16 and out_f32) else None
    _call("conv_umma", (mode, n, h, w, cin, cout), 1, _lib.lib().ptivae_conv_umma, _p(x), _p(w_packed), _p(bias),
          _p(residual), _p(out), _p(out16), _p(part), gn_groups, n, h, w, cin, cout, mode, int(out_f32), res_f32, f16,
          _stream())
    r = (out, part) if gn_groups > 0 else (out,)
    if emit16:
        r = r + (out16,)
    return r if len(r) > 1 else r[0]


def up2x_supported(x: torch.Tensor) -> bool:
    """Shapes/dtypes the halo-resident upsample kernel instantiates (others: conv_umma mode 2)."""
    return x.dtype == F16 and x.shape[-1] in (64, 128)


def up2x_conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, gn_groups: int = 0, emit16: bool = False):
    """conv3x3(nearest_upsample_2x(x)) + bias.  x fp16 NHWC [N,H,W,C] -> fp32 NHWC [N,2H,2W,C]; returns out,
    (out, partials [N,P,G,2]) with gn_groups > 0, and the fp16 copy of out appended when emit16."""
    _need_cuda(x, w_packed, bias)
    if w_packed.dtype != x.dtype:
        raise _lib.PtivaeError("activation and packed-weight operand dtypes differ")
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=torch.float32)
    out16 = torch.empty(out.shape, device=x.device, dtype=x.dtype) if emit16 else None
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_up2x_conv3x3_parts(h, w), gn_groups, 2), device=x.device,
                           dtype=torch.float32)
    _call("up2x_conv3x3", (n, h, w, c, int(emit16)), 1, _lib.lib().ptivae_up2x_conv3x3, _p(x), _p(w_packed), _p(bias),
          _p(out), _p(out16), _p(part), gn_groups, n, h, w, c, _op16(x), _stream())
    r = (out, part) if gn_groups > 0 else (out,)
    if emit16:
        r = r + (out16,)
    return r if len(r) > 1 else r[0]


FUSED_IMPL = 0   # 0 auto | 1 register-staged conv_fused.cu | 2 conv_tma.cu | 3 conv_tma2.cu (tests force each)


def conv3x3_fused(x: torch.Tensor, scale_shift, silu: bool, w_packed: torch.Tensor, bias: torch.Tensor,
                  residual=None, gn_groups: int = 0, out_f32: bool = False):
    """out = conv3x3(act(x*scale+shift)) + bias (+ residual); x NHWC fp32 or 16-bit, weights 16-bit."""
    _need_cuda(x, w_packed, bias)
    n, h, w, cin = x.shape
    cout = w_packed.shape[1]
    f16 = _op16(w_packed)
    if x.dtype != torch.float32 and x.dtype != w_packed.dtype:
        raise _lib.PtivaeError("16-bit input must use the packed-weight dtype")
    out = torch.empty((n, h, w, cout), device=x.device, dtype=torch.float32 if out_f32 else w_packed.dtype)
    res_f32 = 0
    if residual is not None:
        if residual.shape != out.shape:
            raise _lib.PtivaeError("residual shape mismatch")
        res_f32 = int(residual.dtype == torch.float32)
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_conv3x3_fused_parts(h, w), gn_groups, 2), device=x.device,
                           dtype=torch.float32)
    meta = (n, h, w, cin, cout, x.element_size(), out.element_size(),
            0 if residual is None else residual.element_size())
    _call("conv3x3_fused", meta, 1, _lib.lib().ptivae_conv3x3_fused, _p(x), _fmt(x), _p(scale_shift), int(silu),
          _p(w_packed), _p(bias), _p(residual), res_f32, _p(out), int(out_f32), _p(part), gn_groups, n, h, w, cin,
          cout, f16, FUSED_IMPL, _stream())
    return (out, part) if gn_groups > 0 else out


def fused_sc_supported(dtype: torch.dtype, cin: int, cout: int, sc_cin: int) -> bool:
    """(conv2 Cin, Cout, shortcut Cin) combinations for which conv2 + 1x1 shortcut run as one kernel."""
    return dtype == F16 and (cin, cout, sc_cin) in ((32, 32, 64), (64, 64, 32))


def conv3x3_fused_sc(h: torch.Tensor, scale_shift, silu: bool, w_packed: torch.Tensor, bias: torch.Tensor,
                     sc_x: torch.Tensor, sc_w_packed: torch.Tensor, gn_groups: int = 0):
    """out = conv3x3(act(h*scale+shift)) + conv1x1(sc_x) + bias as ONE kernel (bias = sum of both conv biases).
    h, sc_x fp16 NHWC; returns the fp32 stream tensor, or (out, statistics partials) with gn_groups > 0."""
    _need_cuda(h, w_packed, bias, sc_x, sc_w_packed)
    n, hh, w, cin = h.shape
    cout = w_packed.shape[1]
    if sc_x.shape[:3] != h.shape[:3] or sc_x.dtype != h.dtype or sc_w_packed.shape != (1, cout, sc_x.shape[-1]):
        raise _lib.PtivaeError("shortcut operand / weight shape mismatch")
    out = torch.empty((n, hh, w, cout), device=h.device, dtype=torch.float32)
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_conv3x3_fused_parts(hh, w), gn_groups, 2), device=h.device, dtype=torch.float32)
    meta = (n, hh, w, cin, cout, h.element_size(), 4, 0, sc_x.shape[-1])
    _call("conv3x3_fused_sc", meta, 1, _lib.lib().ptivae_conv3x3_fused_sc, _p(h), _p(scale_shift), int(silu), _p(w_packed),
          _p(bias), _p(sc_x), _p(sc_w_packed), sc_x.shape[-1], _p(out), 1, _p(part), gn_groups, n, hh, w, cin, cout, _op16(h),
          _stream())
    return (out, part) if gn_groups > 0 else out


def _nhwc_dims(x):
    n, c = x.shape[0], x.shape[-1]
    return n, x.numel() // (n * c), c


def gn_stats(x: torch.Tensor, groups: int) -> torch.Tensor:
    """-> partial statistics [N,P,G,2] (deterministic; feed to gn_finalize)."""
    _need_cuda(x)
    n, hw, c = _nhwc_dims(x)
    parts = _lib.lib().ptivae_gn_stats_parts(n, hw, c)
    _lib.check(min(parts, 0), "gn_stats_parts")
    part = torch.empty((n, parts, groups, 2), device=x.device, dtype=torch.float32)
    _call("gn_stats", (n, hw, c, x.element_size()), 1, _lib.lib().ptivae_gn_stats, _p(x), _p(part), n, hw, c, groups,
          _fmt(x), _stream())
    return part


def gn_finalize(part: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, hw: int, eps: float) -> torch.Tensor:
    n, parts, g, _ = part.shape
    c = gamma.numel()
    ss = torch.empty((n, c, 2), device=part.device, dtype=torch.float32)
    _call("gn_finalize", None, 1, _lib.lib().ptivae_gn_finalize, _p(part), _p(gamma), _p(beta), _p(ss), n, hw, c, g,
          parts, float(eps), _stream())
    return ss


def gn_apply(x: torch.Tensor, scale_shift: torch.Tensor, silu: bool, emit_raw: bool = False,
             dtype: torch.dtype | None = None):
    """x NHWC (any storage) -> 16-bit operand y (and, if emit_raw, x rounded to that dtype)."""
    n, hw, c = _nhwc_dims(x)
    dtype = dtype or (x.dtype if x.dtype != torch.float32 else F16)
    y = torch.empty(x.shape, device=x.device, dtype=dtype)
    raw = torch.empty(x.shape, device=x.device, dtype=dtype) if emit_raw else None
    _call("gn_apply", (n, hw, c, x.element_size(), int(emit_raw)), 1, _lib.lib().ptivae_gn_apply, _p(x),
          _p(scale_shift), _p(y), _p(raw), n, hw, c, int(silu), _fmt(x), _op16(y), _stream())
    return (y, raw) if emit_raw else y


def small_cin_stats_supported(cin: int, cout: int, groups: int) -> bool:
    """Shapes for which conv3x3_small_cin also emits GroupNorm statistics partials."""
    vecs = cout // 8
    return (cin == 1 and groups > 0 and cout % 8 == 0 and cout % groups == 0 and vecs <= 32 and vecs & (vecs - 1) == 0
            and 8 % (cout // groups) == 0 and 2 * groups <= 256)


def conv3x3_small_cin(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dtype: torch.dtype = torch.float32,
                      gn_groups: int = 0):
    """fp32 NCHW -> NHWC stored as `dtype` (fp32 stream, or a 16-bit operand).  With gn_groups > 0 (see
    small_cin_stats_supported) returns (out, statistics partials [N,P,G,2] of the stored values)."""
    _need_cuda(x, w, b)
    n, cin, h, wd = x.shape
    cout = w.shape[0]
    out = torch.empty((n, h, wd, cout), device=x.device, dtype=dtype)
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_conv3x3_small_cin_parts(h, wd, cout), gn_groups, 2), device=x.device,
                           dtype=torch.float32)
    _call("conv3x3_small_cin", (n, h, wd, cin, cout, out.element_size()), 1, _lib.lib().ptivae_conv3x3_small_cin, _p(x),
          _p(w), _p(b), _p(out), _p(part), gn_groups, n, h, wd, cin, cout, _fmt(out), _stream())
    return (out, part) if gn_groups > 0 else out


def conv3x3_small_cout(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, scale_shift=None) -> torch.Tensor:
    """NHWC bf16|fp32 (+ fused GroupNorm affine) -> fp32 NCHW."""
    _need_cuda(x, w, b)
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    out = torch.empty((n, cout, h, wd), device=x.device, dtype=torch.float32)
    _call("conv3x3_small_cout", (n, h, wd, cin, cout, x.element_size()), 1, _lib.lib().ptivae_conv3x3_small_cout, _p(x), _p(w), _p(b),
          _p(scale_shift), _p(out), n, h, wd, cin, cout, _fmt(x), _stream())
    return out


def conv1x1_small(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, act: int = 0) -> torch.Tensor:
    _need_cuda(x, w, b)
    n, cin = x.shape[:2]
    hw = x.numel() // (n * cin)
    cout = w.shape[0]
    out = torch.empty((n, cout) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
    _call("conv1x1_small", None, 1, _lib.lib().ptivae_conv1x1_small, _p(x), _p(w), _p(b), _p(out), n, hw, cin, cout, act, _stream())
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """16-bit [B,L,D] x3 -> [B,L,D]; single head, scale D^-0.5.  q, k, v may be channel slices of one fused
    [B,L,3D] projection (any common row stride, unit channel stride)."""
    _need_cuda(q, k, v)
    b, l, d = q.shape
    ld = q.stride(1)
    for t in (q, k, v):
        if t.shape != q.shape or t.stride(2) != 1 or t.stride(1) != ld or t.stride(0) != l * ld or t.dtype != q.dtype:
            raise _lib.PtivaeError("q, k, v must share shape, dtype and a [B, L, D] layout with one common row stride")
    out = torch.empty((b, l, d), device=q.device, dtype=q.dtype)
    _call("attention_fwd", (b, l, d), 1, _lib.lib().ptivae_attention_fwd, _p(q), _p(k), _p(v), _p(out), b, l, d, ld,
          _op16(q), _stream())
    return out


def latent_sample(mu, sigma, eps=None, seed: int = 0, offset: int = 0, rng_dev=None, return_eps: bool = False):
    _need_cuda(mu, sigma)
    z = torch.empty_like(mu)
    eps_out = torch.empty_like(mu) if return_eps else None
    _call("latent_sample", None, 1, _lib.lib().ptivae_latent_sample, _p(mu), _p(sigma), _p(eps), _p(z), _p(eps_out), _p(rng_dev),
                                               mu.numel(), seed & (2**64 - 1), offset & (2**64 - 1), _stream())
    return (z, eps_out) if return_eps else z


def rng_advance(rng_dev: torch.Tensor) -> None:
    _call("rng_advance", None, 1, _lib.lib().ptivae_rng_advance, _p(rng_dev), _stream())


def kl_loss(mu: torch.Tensor, t: torch.Tensor, input_is_logvar: bool = True) -> torch.Tensor:
    _need_cuda(mu, t)
    mu = mu.contiguous().float()
    t = t.contiguous().float()
    n = mu.shape[0]
    ws = torch.empty(n, device=mu.device, dtype=torch.float32)
    out = torch.empty(1, device=mu.device, dtype=torch.float32)
    _call("kl_loss", None, 2, _lib.lib().ptivae_kl_loss, _p(mu), _p(t), _p(ws), _p(out), n, mu.numel() // n, int(input_is_logvar),
                                         _stream())
    return out[0]


def l1l2(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """returns a 2-vector: (mean |a-b|, mean (a-b)^2)."""
    _need_cuda(a, b)
    a = a.contiguous().float()
    b = b.contiguous().float()
    if a.shape != b.shape:
        raise _lib.PtivaeError("l1l2 shape mismatch")
    ws = torch.empty(2 * 1184, device=a.device, dtype=torch.float32)
    out = torch.empty(2, device=a.device, dtype=torch.float32)
    _call("l1l2", None, 2, _lib.lib().ptivae_l1l2, _p(a), _p(b), _p(ws), _p(out), a.numel(), _stream())
    return out


def spatial_mean(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 -> [B,C] (mean over H,W)."""
    _need_cuda(x)
    x = x.detach().contiguous().float()
    b, c = x.shape[:2]
    out = torch.empty((b, c), device=x.device, dtype=torch.float32)
    _call("spatial_mean", None, 1, _lib.lib().ptivae_spatial_mean, _p(x), _p(out), b * c, x.numel() // (b * c), _stream())
    return out


def ar_vae_loss(zbar: torch.Tensor, attrs: torch.Tensor, channel: torch.Tensor, delta: torch.Tensor, pairs=None):
    """zbar [B,C], attrs [L,B], channel int32 [L], delta [L], pairs int32 [P,2] | None -> (per-attr loss [L], counts int32 [L], total [1])."""
    _need_cuda(zbar, attrs, channel, delta)
    b, c = zbar.shape
    l = attrs.shape[0]
    loss = torch.empty(l, device=zbar.device, dtype=torch.float32)
    cnt = torch.empty(l, device=zbar.device, dtype=torch.int32)
    tot = torch.empty(1, device=zbar.device, dtype=torch.float32)
    _call("ar_vae_loss", None, 2, _lib.lib().ptivae_ar_vae_loss, _p(zbar), _p(attrs), _p(channel), _p(delta), _p(pairs),
          0 if pairs is None else pairs.shape[0], b, c, l, _p(loss), _p(cnt), _p(tot), _stream())
    return loss, cnt, tot


_ACTS = {None: 0, "none": 0, "relu": 1, "gelu": 2, "leaky_relu": 3, "elu": 4}


def linear_act(x: torch.Tensor, w: torch.Tensor, b, act: str | None = None) -> torch.Tensor:
    """fp32 [B,I] x [O,I]^T + b, optional activation -> [B,O]."""
    _need_cuda(x, w)
    x = x.detach().contiguous().float()
    w = w.detach().contiguous().float()
    bsz, i = x.shape
    o = w.shape[0]
    y = torch.empty((bsz, o), device=x.device, dtype=torch.float32)
    _call("linear_act", None, 1, _lib.lib().ptivae_linear_act, _p(x), _p(w), _p(None if b is None else b.detach()), _p(y),
          bsz, i, o, _ACTS[act], _stream())
    return y


def eval_metrics(pred: torch.Tensor, target: torch.Tensor, window: torch.Tensor, clamp=(0.0, 1.0), data_range: float = 1.0,
                 k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """fp32 NCHW pred/target -> [B,4] = per-sample (mse, mae, psnr, ssim); `clamp` = (lo, hi) applied to both images
    first, or None.  window: fp32 taps of the separable SSIM window (odd length <= 15)."""
    _need_cuda(pred, target, window)
    if pred.shape != target.shape or pred.dim() != 4 or pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise _lib.PtivaeError("eval_metrics needs two fp32 NCHW tensors of the same shape")
    pred, target, window = pred.contiguous(), target.contiguous(), window.contiguous().float()
    b, c, h, w = pred.shape
    out = torch.empty((b, 4), device=pred.device, dtype=torch.float32)
    nbytes = _lib.lib().ptivae_eval_metrics_workspace(b, c, h, w)
    if nbytes < 0:
        _lib.check(nbytes, "eval_metrics_workspace")
    ws = torch.empty(nbytes, device=pred.device, dtype=torch.uint8)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    _call("eval_metrics", (b, c, h, w), 2, _lib.lib().ptivae_eval_metrics, _p(pred), _p(target), _p(window), window.numel(),
          _p(out), _p(ws), b, c, h, w, int(clamp is not None), float(lo), float(hi), float(data_range), float(k1), float(k2),
          _stream())
    return out


def local_normalize(x: torch.Tensor, return_stats: bool = False):
    """Batch of fp32 images [B, ...] -> z-score over each image's non-zero pixels, zeros stay zero."""
    _need_cuda(x)
    if x.dtype != torch.float32:
        raise _lib.PtivaeError("local_normalize needs an fp32 tensor")
    x = x.contiguous()
    b = x.shape[0]
    per = x.numel() // b
    out = torch.empty_like(x)
    stats = torch.empty((b, 2), device=x.device, dtype=torch.float32) if return_stats else None
    ws = torch.empty(_lib.lib().ptivae_local_normalize_workspace(b), device=x.device, dtype=torch.uint8)
    _call("local_normalize", (b, per), 2, _lib.lib().ptivae_local_normalize, _p(x), _p(out), _p(stats), _p(ws), b, per,
          _stream())
    return (out, stats) if return_stats else out
