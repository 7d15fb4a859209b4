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
    """Every tensor handed to the C ABI must live on the CURRENT CUDA device: the kernels are launched on
    torch.cuda.current_stream() of that device with raw pointers (the model-level entry points switch the current
    device to the model's device for the duration of a call, so a model on cuda:1 works with cuda:0 current)."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.PtivaeError("ptivae ops need CUDA tensors (there is no CPU path)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.PtivaeError(f"tensor on {t.device} but the current CUDA device is cuda:{cur}: wrap the call in "
                                   "torch.cuda.device(tensor.device) (kernels launch on the current device's stream)")


def pack_conv_weight(w: torch.Tensor, mode: int = 0, dtype: torch.dtype = F16) -> torch.Tensor:
    """fp32 [Cout,Cin,k,k] (or Linear [out,in]) -> 16-bit [T,Cout,Cin]; mode | 4: transposed slabs [T,Cin,Cout]
    (the operand of the data-gradient convolutions, conv_umma modes 4/5/6 and the transposed 1x1)."""
    _need_cuda(w)
    if w.dim() == 2:
        w = w[:, :, None, None]
    w = w.detach().contiguous().float()
    cout, cin, k, _ = w.shape
    t = 16 if (mode & 3) == 2 else k * k
    out = torch.empty((t, cin, cout) if mode & 4 else (t, cout, cin), device=w.device, dtype=dtype)
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
    if w_packed.shape[2] != cin or bias.numel() != cout or h <= 0 or w <= 0:
        raise _lib.PtivaeError(f"conv_umma: activation {tuple(x.shape)} / packed weight {tuple(w_packed.shape)} / bias "
                               f"{bias.numel()} mismatch")
    ho, wo = (h // 2, w // 2) if mode in (1, 6) else ((2 * h, 2 * w) if mode in (2, 5) else (h, w))
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


def up2x_conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, gn_groups: int = 0, emit16: bool = False,
                 out_f32: bool = True):
    """conv3x3(nearest_upsample_2x(x)) + bias.  x fp16 NHWC [N,H,W,C] -> fp32 NHWC [N,2H,2W,C]; returns out,
    (out, partials [N,P,G,2]) with gn_groups > 0, and the fp16 copy of out appended when emit16.  out_f32=False
    (16-bit residual stream): the output itself is fp16 (statistics of the stored values), no fp32 tensor is written."""
    _need_cuda(x, w_packed, bias)
    if w_packed.dtype != x.dtype:
        raise _lib.PtivaeError("activation and packed-weight operand dtypes differ")
    n, h, w, c = x.shape
    if not out_f32:
        out, out16, emit16 = None, torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=x.dtype), False
    else:
        out = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=torch.float32)
        out16 = torch.empty(out.shape, device=x.device, dtype=x.dtype) if emit16 else None
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_up2x_conv3x3_parts(h, w), gn_groups, 2), device=x.device,
                           dtype=torch.float32)
    _call("up2x_conv3x3", (n, h, w, c, int(emit16) if out_f32 else 2), 1, _lib.lib().ptivae_up2x_conv3x3, _p(x), _p(w_packed),
          _p(bias), _p(out), _p(out16), _p(part), gn_groups, n, h, w, c, _op16(x), _stream())
    if not out_f32:
        out = out16
    r = (out, part) if gn_groups > 0 else (out,)
    if emit16:
        r = r + (out16,)
    return r if len(r) > 1 else r[0]


FUSED_IMPL = 0   # 0 auto | 1 register-staged conv_fused.cu | 2 conv_tma.cu | 3 conv_tma2.cu | 4 conv_band.cu | 5 conv_pair.cu (tests force each)


def conv3x3_fused(x: torch.Tensor, scale_shift, silu: bool, w_packed: torch.Tensor, bias: torch.Tensor,
                  residual=None, gn_groups: int = 0, out_f32: bool = False):
    """out = conv3x3(act(x*scale+shift)) + bias (+ residual); x NHWC fp32 or 16-bit, weights 16-bit."""
    _need_cuda(x, w_packed, bias)
    n, h, w, cin = x.shape
    cout = w_packed.shape[1]
    f16 = _op16(w_packed)
    if w_packed.shape[2] != cin or bias.numel() != cout or h <= 0 or w <= 0:
        raise _lib.PtivaeError(f"conv3x3_fused: activation {tuple(x.shape)} / packed weight {tuple(w_packed.shape)} / bias "
                               f"{bias.numel()} mismatch")
    if x.dtype != torch.float32 and x.dtype != w_packed.dtype:
        raise _lib.PtivaeError("16-bit input must use the packed-weight dtype")
    out = torch.empty((n, h, w, cout), device=x.device, dtype=torch.float32 if out_f32 else w_packed.dtype)
    res_f32 = 0
    if residual is not None:
        if residual.shape != out.shape:
            raise _lib.PtivaeError("residual shape mismatch")
        res_f32 = int(residual.dtype == torch.float32)
        if not res_f32 and residual.dtype != w_packed.dtype:
            raise _lib.PtivaeError("a 16-bit residual must use the operand dtype")
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


def set_chained_launch(mask: int) -> int:
    """Programmatic dependent launches for the forward chain (ptivae.h): bit 0 the tensor-core kernels (default), bit 1 the
    small kernels; returns the previous mask."""
    return int(_lib.lib().ptivae_set_chained_launch(int(mask)))


def conv3x3_fused_supported(x_dtype, res_dtype, out_f32: bool, cin: int, cout: int, op_dtype) -> bool:
    """Whether conv3x3_fused has a kernel for this combination (widths 32/64/128 always; 256 on the 16-bit stream)."""
    in_fmt = _FMT.get(x_dtype, -1)
    res_kind = 0 if res_dtype is None else (1 if res_dtype == torch.float32 else 2)
    return _lib.lib().ptivae_conv3x3_fused_query(in_fmt, res_kind, int(out_f32), cin, cout, int(op_dtype == F16)) == 0


def fused_sc_supported(dtype: torch.dtype, cin: int, cout: int, sc_cin: int) -> bool:
    """(conv2 Cin, Cout, shortcut Cin) combinations for which conv2 + 1x1 shortcut run as one kernel."""
    return dtype == F16 and (cin, cout, sc_cin) in ((32, 32, 64), (64, 64, 32))


def conv3x3_fused_sc(h: torch.Tensor, scale_shift, silu: bool, w_packed: torch.Tensor, bias: torch.Tensor,
                     sc_x: torch.Tensor, sc_w_packed: torch.Tensor, gn_groups: int = 0, out_f32: bool = True):
    """out = conv3x3(act(h*scale+shift)) + conv1x1(sc_x) + bias as ONE kernel (bias = sum of both conv biases).
    h, sc_x fp16 NHWC; returns the fp32 stream tensor, or (out, statistics partials) with gn_groups > 0."""
    _need_cuda(h, w_packed, bias, sc_x, sc_w_packed)
    n, hh, w, cin = h.shape
    cout = w_packed.shape[1]
    if sc_x.shape[:3] != h.shape[:3] or sc_x.dtype != h.dtype or sc_w_packed.shape != (1, cout, sc_x.shape[-1]):
        raise _lib.PtivaeError("shortcut operand / weight shape mismatch")
    out = torch.empty((n, hh, w, cout), device=h.device, dtype=torch.float32 if out_f32 else h.dtype)
    part = None
    if gn_groups > 0:
        part = torch.empty((n, _lib.lib().ptivae_conv3x3_fused_parts(hh, w), gn_groups, 2), device=h.device, dtype=torch.float32)
    meta = (n, hh, w, cin, cout, h.element_size(), out.element_size(), 0, sc_x.shape[-1])
    _call("conv3x3_fused_sc", meta, 1, _lib.lib().ptivae_conv3x3_fused_sc, _p(h), _p(scale_shift), int(silu), _p(w_packed),
          _p(bias), _p(sc_x), _p(sc_w_packed), sc_x.shape[-1], _p(out), int(out_f32), _p(part), gn_groups, n, hh, w, cin, cout, _op16(h),
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


def gn_finalize(part: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, hw: int, eps: float,
                return_mean_rstd: bool = False, range_flag: torch.Tensor | None = None):
    """partials [N,P,G,2] -> scale/shift [N,C,2] (and, for the backward pass, (mean, rstd) [N,G,2]).  range_flag (int32 [1],
    optional): set to 1 when a partial's sum of squares reaches 65504^2 (fp16 range check, see include/ptivae.h)."""
    n, parts, g, _ = part.shape
    c = gamma.numel()
    ss = torch.empty((n, c, 2), device=part.device, dtype=torch.float32)
    mr = torch.empty((n, g, 2), device=part.device, dtype=torch.float32) if return_mean_rstd else None
    if range_flag is None:
        _call("gn_finalize", None, 1, _lib.lib().ptivae_gn_finalize, _p(part), _p(gamma), _p(beta), _p(ss), _p(mr), n, hw, c, g,
              parts, float(eps), _stream())
    else:
        _call("gn_finalize", None, 1, _lib.lib().ptivae_gn_finalize_checked, _p(part), _p(gamma), _p(beta), _p(ss), _p(mr), n, hw,
              c, g, parts, float(eps), _p(range_flag), _stream())
    return (ss, mr) if return_mean_rstd else ss


def range_check(part: torch.Tensor, range_flag: torch.Tensor) -> None:
    """fp16 range check on the statistics partials [N,P,G,2] of a tensor that no GroupNorm consumes."""
    _need_cuda(part, range_flag)
    _call("range_check", None, 1, _lib.lib().ptivae_range_check, _p(part), part.numel() // 2, _p(range_flag), _stream())


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
    if tuple(w.shape[1:]) != (cin, 3, 3) or b.numel() != cout or h <= 0 or wd <= 0:
        raise _lib.PtivaeError(f"conv3x3_small_cin: input {tuple(x.shape)} does not match weight {tuple(w.shape)} / bias {b.numel()}")
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
    if tuple(w.shape[1:]) != (cin, 3, 3) or b.numel() != cout or (scale_shift is not None and tuple(scale_shift.shape) != (n, cin, 2)):
        raise _lib.PtivaeError(f"conv3x3_small_cout: input {tuple(x.shape)} does not match weight {tuple(w.shape)} / bias {b.numel()}")
    out = torch.empty((n, cout, h, wd), device=x.device, dtype=torch.float32)
    _call("conv3x3_small_cout", (n, h, wd, cin, cout, x.element_size()), 1, _lib.lib().ptivae_conv3x3_small_cout, _p(x), _p(w), _p(b),
          _p(scale_shift), _p(out), n, h, wd, cin, cout, _fmt(x), _stream())
    return out


def conv1x1_small(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, act: int = 0) -> torch.Tensor:
    _need_cuda(x, w, b)
    n, cin = x.shape[:2]
    hw = x.numel() // (n * cin)
    cout = w.shape[0]
    if w.numel() != cout * cin or b.numel() != cout:
        raise _lib.PtivaeError(f"conv1x1_small: input {tuple(x.shape)} does not match weight {tuple(w.shape)} / bias {b.numel()}")
    out = torch.empty((n, cout) + tuple(x.shape[2:]), device=x.device, dtype=torch.float32)
    _call("conv1x1_small", None, 1, _lib.lib().ptivae_conv1x1_small, _p(x), _p(w), _p(b), _p(out), n, hw, cin, cout, act, _stream())
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, return_lse: bool = False):
    """16-bit [B,L,D] x3 -> [B,L,D]; single head, scale D^-0.5.  q, k, v may be channel slices of one fused
    [B,L,3D] projection (any common row stride, unit channel stride).  return_lse: also the fp32 [B,L] log2-domain
    log-sum-exp of the scaled score rows (what attention_bwd needs)."""
    _need_cuda(q, k, v)
    b, l, d = q.shape
    ld = q.stride(1)
    for t in (q, k, v):
        if t.shape != q.shape or t.stride(2) != 1 or t.stride(1) != ld or t.stride(0) != l * ld or t.dtype != q.dtype:
            raise _lib.PtivaeError("q, k, v must share shape, dtype and a [B, L, D] layout with one common row stride")
    out = torch.empty((b, l, d), device=q.device, dtype=q.dtype)
    lse = torch.empty((b, l), device=q.device, dtype=torch.float32) if return_lse else None
    _call("attention_fwd", (b, l, d), 1, _lib.lib().ptivae_attention_fwd, _p(q), _p(k), _p(v), _p(out), _p(lse), b, l, d, ld,
          _op16(q), _stream())
    return (out, lse) if return_lse else out


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
    n = a.numel()
    if n == 0:
        raise _lib.PtivaeError("l1l2 of empty tensors")
    ws = torch.empty(2 * 1184, device=a.device, dtype=torch.float32)
    out = torch.empty(2, device=a.device, dtype=torch.float32)
    if n % 4 != 0 or a.data_ptr() % 16 != 0 or b.data_ptr() % 16 != 0:
        # the kernel reads float4: any other shape / alignment (nn.L1Loss accepts all) goes through zero-padded copies
        # (|0 - 0| adds nothing to either sum) and the means are rescaled to the true element count
        m = (n + 3) // 4 * 4
        ap, bp = torch.zeros(m, device=a.device, dtype=torch.float32), torch.zeros(m, device=a.device, dtype=torch.float32)
        ap[:n].copy_(a.reshape(-1))
        bp[:n].copy_(b.reshape(-1))
        _call("l1l2", None, 2, _lib.lib().ptivae_l1l2, _p(ap), _p(bp), _p(ws), _p(out), m, _stream())
        return out * (m / n)
    _call("l1l2", None, 2, _lib.lib().ptivae_l1l2, _p(a), _p(b), _p(ws), _p(out), n, _stream())
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


def ar_vae_loss_bwd(zbar, attrs, channel, delta, pairs, count, g_total=None, g_attr=None) -> torch.Tensor:
    """Gradient of ar_vae_loss w.r.t. zbar [B,C]; g_total: device scalar dL/d(total), g_attr: [L] dL/d(per-attribute loss)."""
    _need_cuda(zbar, attrs, channel, delta, count)
    b, c = zbar.shape
    l = attrs.shape[0]
    dz = torch.empty((b, c), device=zbar.device, dtype=torch.float32)
    _call("ar_vae_loss_bwd", None, 1, _lib.lib().ptivae_ar_vae_loss_bwd, _p(zbar), _p(attrs), _p(channel), _p(delta), _p(pairs),
          0 if pairs is None else pairs.shape[0], b, c, l, _p(count), _p(g_total), _p(g_attr), _p(dz), _stream())
    return dz


def spatial_mean_bwd(dmean: torch.Tensor, shape) -> torch.Tensor:
    """[B,C] gradient of spatial_mean -> [B,C,H,W] (each pixel gets dmean / (H*W))."""
    _need_cuda(dmean)
    dmean = dmean.detach().contiguous().float()
    dx = torch.empty(tuple(shape), device=dmean.device, dtype=torch.float32)
    b, c = shape[0], shape[1]
    _call("spatial_mean_bwd", None, 1, _lib.lib().ptivae_spatial_mean_bwd, _p(dmean), _p(dx), b * c, dx.numel() // (b * c), _stream())
    return dx


_ACTS = {None: 0, "none": 0, "relu": 1, "gelu": 2, "leaky_relu": 3, "elu": 4}


def linear_act(x: torch.Tensor, w: torch.Tensor, b, act: str | None = None) -> torch.Tensor:
    """fp32 [B,I] x [O,I]^T + b, optional activation -> [B,O]."""
    _need_cuda(x, w)
    x = x.detach().contiguous().float()
    w = w.detach().contiguous().float()
    bsz, i = x.shape
    o = w.shape[0]
    y = torch.empty((bsz, o), device=x.device, dtype=torch.float32)
    bias = None if b is None else b.detach().contiguous().float()
    if w.shape[1] != i or (bias is not None and bias.numel() != o):
        raise _lib.PtivaeError(f"linear_act: input {tuple(x.shape)} does not match weight {tuple(w.shape)}")
    _call("linear_act", None, 1, _lib.lib().ptivae_linear_act, _p(x), _p(w), _p(bias), _p(y),
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


def resize_area(x: torch.Tensor, size) -> torch.Tensor:
    """[B,H,W] uint8 | uint16 (stored as int16 / uint16) | float32 -> fp32 [B,Ho,Wo], area interpolation (MONAI Resize default)."""
    _need_cuda(x)
    fmt = {torch.uint8: 0, torch.int16: 1, torch.float32: 2}.get(x.dtype)
    if fmt is None and hasattr(torch, "uint16") and x.dtype == torch.uint16:
        fmt = 1
    if fmt is None or x.dim() != 3:
        raise _lib.PtivaeError("resize_area takes a [B, H, W] uint8 / uint16 / float32 tensor")
    x = x.contiguous()
    b, h, w = x.shape
    ho, wo = int(size[0]), int(size[1])
    out = torch.empty((b, ho, wo), device=x.device, dtype=torch.float32)
    _call("resize_area", None, 1, _lib.lib().ptivae_resize_area, _p(x), fmt, _p(out), b, h, w, ho, wo, _stream())
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


# ---------------------------------------------------------------------------------------------------------------
# backward pass
# ---------------------------------------------------------------------------------------------------------------
def wgrad(dy: torch.Tensor, x: torch.Tensor, mode: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Weight gradient of a conv on tensor cores.  dy: 16-bit NHWC gradient of the conv output, x: 16-bit NHWC conv input
    (as the forward GEMM read it).  mode as conv_umma (0 3x3 | 1 pad+3x3 s2 | 2 up2x+3x3 | 3 1x1).
    -> fp32 [Cout,Cin,3,3] (or [Cout,Cin,1,1]), written into `out` when given (a view of the flat gradient buffer)."""
    _need_cuda(dy, x)
    n, h, w, cb = x.shape
    ca = dy.shape[-1]
    exp = (n, h // 2, w // 2) if mode == 1 else ((n, 2 * h, 2 * w) if mode == 2 else (n, h, w))
    if tuple(dy.shape[:3]) != exp:
        raise _lib.PtivaeError(f"wgrad: dy {tuple(dy.shape)} does not match x {tuple(x.shape)} for mode {mode}")
    k = 1 if mode == 3 else 3      # (mode 4: the 3x3 s1 gradient through the generic per-tap kernel)
    if out is None:
        out = torch.empty((ca, cb, k, k), device=x.device, dtype=torch.float32)
    elif out.numel() != ca * cb * k * k or out.dtype != torch.float32 or not out.is_contiguous():
        raise _lib.PtivaeError("wgrad: bad output buffer")
    nbytes = _lib.lib().ptivae_wgrad_workspace(n, h, w, ca, cb, mode)
    if nbytes < 0:
        _lib.check(int(nbytes), "wgrad_workspace")
    ws = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
    if dy.dtype != x.dtype:
        raise _lib.PtivaeError("wgrad: both operands of one tcgen05 MMA must share a 16-bit format")
    _call("wgrad", (mode, n, h, w, ca, cb), 2, _lib.lib().ptivae_wgrad, _p(dy), _p(x), _p(ws), _p(out), n, h, w, ca, cb,
          mode, _op16(dy), _stream())
    return out


def bgemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, a_mn: bool, b_mn: bool, epi: int = 0, alpha: float = 1.0,
          rowv: torch.Tensor | None = None, aux: torch.Tensor | None = None, k: int | None = None) -> torch.Tensor:
    """out[b,m,n] = epi(sum_k A(m,k)*B(n,k)) for 3-D 16-bit views with unit inner stride.  a is [B,M,K] (a_mn False) or
    [B,K,M] (a_mn True); likewise b with N; `k` overrides the contraction length (padded row strides)."""
    _need_cuda(a, b, out)
    for t in (a, b, out) + ((aux,) if aux is not None else ()):
        if t.dim() != 3 or t.stride(2) != 1:
            raise _lib.PtivaeError("bgemm operands must be 3-D with unit inner stride")
    bsz, m, n = out.shape
    if a.dtype != b.dtype:
        raise _lib.PtivaeError("bgemm: both operands of one tcgen05 MMA must share a 16-bit format")
    kk = k if k is not None else (a.shape[1] if a_mn else a.shape[2])
    _call("bgemm", (bsz, m, n, kk, int(a_mn), int(b_mn), epi), 1, _lib.lib().ptivae_bgemm, _p(a), _p(b), _p(out), bsz, m, n,
          kk, a.stride(1), a.stride(0), int(a_mn), _op16(a), b.stride(1), b.stride(0), int(b_mn), _op16(b), out.stride(1),
          out.stride(0), _op16(out), epi, float(alpha), _p(rowv), _p(aux), 0 if aux is None else aux.stride(1),
          0 if aux is None else aux.stride(0), 0 if aux is None else _op16(aux), _stream())
    return out


def rowdot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """16-bit [B,L,D] x2 (unit inner stride, dense rows) -> fp32 [B,L] row dot products."""
    _need_cuda(a, b)
    bsz, l, d = a.shape
    for t in (a, b):
        if t.stride(2) != 1 or t.stride(0) != l * t.stride(1):
            raise _lib.PtivaeError("rowdot needs [B,L,D] views with one row stride")
    out = torch.empty((bsz, l), device=a.device, dtype=torch.float32)
    _call("rowdot", (bsz, l, d), 1, _lib.lib().ptivae_rowdot, _p(a), _p(b), _p(out), bsz * l, d, a.stride(1), b.stride(1),
          _op16(a), _op16(b), _stream())
    return out


def cast16(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Storage conversion to a 16-bit operand format (contiguous input, numel % 8 == 0)."""
    _need_cuda(x)
    if x.dtype == dtype:
        return x
    if not x.is_contiguous():
        raise _lib.PtivaeError("cast16 needs a contiguous tensor")
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    _call("cast16", (x.numel(),), 1, _lib.lib().ptivae_cast16, _p(x), _p(out), x.numel(), _fmt(x), _op16(out), _stream())
    return out


def attention_bwd(q, k, v, o, lse, d_o, dqkv: torch.Tensor) -> None:
    """Backward of ops.attention.  q,k,v: bf16 [B,L,D] views (row stride ld) of the forward's projections; o: forward
    output (any 16-bit format); lse: from the forward; d_o: bf16 [B,L,D] gradient of o.  Writes dq|dk|dv into the three
    channel slices of dqkv bf16 [B,L,3D]."""
    bsz, l, d = q.shape
    lp = (l + 7) // 8 * 8
    scale = float(d) ** -0.5
    drow = rowdot(d_o, o)
    p = torch.empty((bsz, l, lp), device=q.device, dtype=torch.bfloat16)[:, :, :l]
    ds = torch.empty((bsz, l, lp), device=q.device, dtype=torch.bfloat16)[:, :, :l]
    bgemm(q, k, p, False, False, epi=1, alpha=scale * 1.4426950408889634, rowv=lse)           # P = exp2(QK^T c - lse)
    bgemm(p, d_o, dqkv[:, :, 2 * d:], True, True)                                            # dV = P^T dO
    bgemm(d_o, v, ds, False, False, epi=2, alpha=scale, rowv=drow, aux=p)                    # dS = P o (dO V^T - D) * scale
    bgemm(ds, k, dqkv[:, :, :d], False, True)                                                # dQ = dS K
    bgemm(ds, q, dqkv[:, :, d:2 * d], True, True)                                            # dK = dS^T Q


def gn_bwd(x: torch.Tensor, da: torch.Tensor, ss: torch.Tensor, mr: torch.Tensor, gamma: torch.Tensor, silu: bool,
           dgamma: torch.Tensor, dbeta: torch.Tensor, residual=None, want32: bool = True, want16: bool = True,
           want_act: bool = False, colsum_out: torch.Tensor | None = None):
    """Backward of act(GroupNorm(x)).  -> (dx fp32 | None, dx bf16 | None[, act bf16]); dgamma/dbeta (fp32 [C]) are
    overwritten.  want_act: also return act(GroupNorm(x)) in bf16 (the weight-gradient operand); colsum_out: fp32 [C]
    that receives the per-channel sum of dx (the bias gradient of the conv that produced x)."""
    _need_cuda(x, da, ss, mr, gamma)
    n, hw, c = _nhwc_dims(x)
    g = mr.shape[1]
    if da.shape != x.shape or (residual is not None and residual.shape != x.shape):
        raise _lib.PtivaeError("gn_bwd shape mismatch")
    ws = torch.empty(_lib.lib().ptivae_gn_bwd_workspace(n, hw, c), device=x.device, dtype=torch.float32)
    coef = torch.empty((n, c, 2), device=x.device, dtype=torch.float32)
    dx32 = torch.empty(x.shape, device=x.device, dtype=torch.float32) if want32 else None
    dx16 = torch.empty(x.shape, device=x.device, dtype=BF16) if want16 else None
    act = torch.empty(x.shape, device=x.device, dtype=BF16) if want_act else None
    _call("gn_bwd", (n, hw, c, x.element_size(), int(silu)), 4 + int(colsum_out is not None), _lib.lib().ptivae_gn_bwd, _p(x),
          _fmt(x), _p(da), _fmt(da), _p(ss), _p(mr), _p(gamma), _p(residual), 0 if residual is None else _fmt(residual),
          _p(dx32), _p(dx16), _p(dgamma), _p(dbeta), _p(act), _p(colsum_out), _p(coef), _p(ws), n, hw, c, g, int(silu),
          _stream())
    return (dx32, dx16, act) if want_act else (dx32, dx16)


def colsum(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """Per-channel sum of an NHWC tensor (bias gradient) -> fp32 [C]."""
    _need_cuda(x)
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty(c, device=x.device, dtype=torch.float32)
    ws = torch.empty(_lib.lib().ptivae_colsum_blocks(rows) * c, device=x.device, dtype=torch.float32)
    _call("colsum", (rows, c), 2, _lib.lib().ptivae_colsum, _p(x), _p(out), _p(ws), rows, c, _fmt(x), _stream())
    return out


def thin_wgrad(thin: torch.Tensor, wide: torch.Tensor, wide_is_input: bool, dw: torch.Tensor, db=None, scale_shift=None):
    """Weight gradient of a thin-end 3x3 conv (see include/ptivae.h).  thin fp32 NCHW, wide NHWC."""
    _need_cuda(thin, wide, dw)
    n, ct, h, w = thin.shape
    c = wide.shape[-1]
    if tuple(wide.shape[:3]) != (n, h, w) or dw.numel() != ct * c * 9:
        raise _lib.PtivaeError("thin_wgrad shape mismatch")
    ws = torch.empty(_lib.lib().ptivae_thin_wgrad_workspace(n, h, w, c, ct) // 4, device=thin.device, dtype=torch.float32)
    _call("thin_wgrad", (n, h, w, c, ct), 2, _lib.lib().ptivae_thin_wgrad, _p(thin), _p(wide), _p(scale_shift), _p(dw),
          _p(db), _p(ws), n, h, w, c, ct, _fmt(wide), int(wide_is_input), _stream())
    return dw


def latent_bwd(dzq, dmu_ext, dsig_ext, eps, h, mu, sigma, wp, wm, ws_, bs):
    """-> (dh, dmu, dlv, z): see include/ptivae.h ptivae_latent_bwd.  All fp32 NCHW [N,L,h,w]."""
    _need_cuda(dzq, eps, h, mu, sigma)
    n, l = h.shape[:2]
    hw = h.numel() // (n * l)
    dh, dmu, dlv, z = (torch.empty_like(h) for _ in range(4))
    _call("latent_bwd", None, 1, _lib.lib().ptivae_latent_bwd, _p(dzq), _p(dmu_ext), _p(dsig_ext), _p(eps), _p(h), _p(mu),
          _p(sigma), _p(wp), _p(wm), _p(ws_), _p(bs), _p(dh), _p(dmu), _p(dlv), _p(z), n, hw, l, _stream())
    return dh, dmu, dlv, z


def outer_reduce(a: torch.Tensor, b: torch.Tensor, dw: torch.Tensor, db=None):
    """dw[i,j] = sum_{n,p} a[n,i,p] b[n,j,p]; db[i] = sum a[n,i,p] (gradients of a 1x1 latent conv)."""
    n, i = a.shape[:2]
    j = b.shape[1]
    hw = a.numel() // (n * i)
    _call("outer_reduce", None, 1, _lib.lib().ptivae_outer_reduce, _p(a), _p(b), _p(dw), _p(db), n, i, j, hw, _stream())
    return dw


def l1l2_bwd(a: torch.Tensor, b: torch.Tensor, gout: torch.Tensor) -> torch.Tensor:
    """gout: fp32 device 2-vector (dL/d l1, dL/d l2) -> dL/da."""
    d = torch.empty_like(a)
    _call("l1l2_bwd", None, 1, _lib.lib().ptivae_l1l2_bwd, _p(a), _p(b), _p(gout), _p(d), a.numel(), _stream())
    return d


def kl_bwd(mu: torch.Tensor, t: torch.Tensor, gout: torch.Tensor, input_is_logvar: bool = True):
    n = mu.shape[0]
    dmu, dt = torch.empty_like(mu), torch.empty_like(t)
    _call("kl_bwd", None, 1, _lib.lib().ptivae_kl_bwd, _p(mu), _p(t), _p(gout), _p(dmu), _p(dt), n, mu.numel() // n,
          int(input_is_logvar), _stream())
    return dmu, dt


def adam(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step_dev: torch.Tensor, lr: float,
         betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0, advance: bool = True) -> None:
    """torch.optim.Adam (defaults) over flat fp32 buffers, in place; step_dev = device float with the 1-based step."""
    _need_cuda(p, g, m, v, step_dev)
    _call("adam", None, 2 if advance else 1, _lib.lib().ptivae_adam, _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr),
          float(betas[0]), float(betas[1]), float(eps), float(grad_scale), _p(step_dev), int(advance), _stream())


def build_pack_table(recipes):
    """recipes: list of (src fp32 weight [Cout,Cin,k,k] | [out,in], dst 16-bit tensor, dst element offset, st_r, mode)
    -> (device table, n, total) for pack_many.  mode as pack_conv_weight (0 plain | 2 up2x 16-slab | +4 transposed)."""
    import numpy as np
    dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("start", "<i8"), ("Cout", "<i4"), ("Cin", "<i4"), ("ksq", "<i4"),
                   ("T", "<i4"), ("st_t", "<i8"), ("st_r", "<i4"), ("f16", "<i4"), ("transpose", "<i4"), ("up2x", "<i4")])
    if dt.itemsize != _lib.lib().ptivae_pack_desc_bytes():
        raise _lib.PtivaeError("pack descriptor layout mismatch between ops.py and latent_loss.cu")
    tab = np.zeros(len(recipes), dtype=dt)
    start = 0
    dev = None
    for i, (w, dst, off, st_r, mode) in enumerate(recipes):
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise _lib.PtivaeError("pack_many sources must be contiguous fp32 master weights")
        cout, cin = w.shape[0], w.shape[1]
        ksq = w.numel() // (cout * cin)
        up = (mode & 3) == 2
        t = 16 if up else ksq
        tab[i] = (w.data_ptr(), dst.data_ptr() + off * 2, start, cout, cin, ksq, t, dst.shape[1] * dst.shape[2], st_r,
                  int(dst.dtype == F16), int(bool(mode & 4)), int(up))
        start += t * cout * cin
        dev = w.device
    table = torch.from_numpy(tab.view(np.uint8).reshape(-1).copy()).to(dev)
    return table, len(recipes), start


def pack_many(table: torch.Tensor, n: int, total: int) -> None:
    _call("pack_many", None, 1, _lib.lib().ptivae_pack_many, _p(table), n, total, _stream())
