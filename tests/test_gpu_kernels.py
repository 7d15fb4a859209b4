"""Per-kernel parity on the GPU: every C-ABI entry point vs. a plain PyTorch fp32 statement of the same
operator on identical (16-bit-rounded) inputs, for both operand formats (fp16 = default, bf16).
Tolerances are written next to each check:
  * 16-bit-output kernels: the result must sit within ~1 ulp of the fp32 reference
    (bf16: rel-L2 <= 3e-3, max-abs <= 2^-7 * max|ref|; fp16: 8x tighter);
  * fp32-accumulate / fp32-output kernels: max-abs <= 1e-4 relative to the reference scale.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _check_bf16(out, ref, what, rel=3e-3, ulp=2.0 ** -7):
    if out.dtype == torch.float16:  # 11-bit significand instead of 8
        rel, ulp = rel / 8, ulp / 8
    out, ref = out.float(), ref.float()
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    r = _rel_l2(out, ref)
    mx = float((out - ref).abs().max())
    scale = float(ref.abs().max())
    assert r <= rel, f"{what}: rel-L2 {r:.3e} > {rel}"
    assert mx <= ulp * scale + 1e-6, f"{what}: max-abs {mx:.3e} vs scale {scale:.3e}"


DT = torch.float16  # operand format under test; the fixture below switches it


@pytest.fixture(params=[torch.float16, torch.bfloat16], ids=["fp16", "bf16"], autouse=True)
def _operand_dtype(request):
    global DT
    DT = request.param
    yield
    DT = torch.float16


def _rand_act(n, h, w, c, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(n, h, w, c, generator=g).to(DEV).to(DT)


def _rand_conv(cout, cin, k, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(DEV)
    b = (torch.randn(cout, generator=g) * 0.1).to(DEV)
    return w.to(DT).float(), b  # weights exactly representable in the operand format


def _ref_conv(x_nhwc, w, b, mode):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    if mode == 0:
        y = F.conv2d(x, w, b, padding=1)
    elif mode == 1:
        y = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    elif mode == 2:
        y = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    else:
        y = F.conv2d(x, w, b)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.cuda.synchronize()


def test_pack_weights(b200):
    w, _ = _rand_conv(64, 32, 3, 1)
    p = b200.ops.pack_conv_weight(w, 0, DT)
    assert p.dtype == DT
    ref = w.permute(2, 3, 0, 1).reshape(9, 64, 32)
    assert torch.equal(p.float(), ref)
    p2 = b200.ops.pack_conv_weight(w, 2, DT).float().view(2, 2, 2, 2, 64, 32)
    rows = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    s = sum(w[:, :, ky, kx] for ky in rows[py][ty] for kx in rows[px][tx])
                    assert torch.equal(p2[py, px, ty, tx], s.to(DT).float())
    lin = torch.randn(128, 128, device=DEV).to(DT).float()
    assert torch.equal(b200.ops.pack_conv_weight(lin, 0, DT).float()[0], lin)


CONV_CASES = [
    # (mode, N, H, W, Cin, Cout)
    (0, 2, 32, 32, 128, 128),
    (0, 1, 16, 16, 64, 64),
    (0, 2, 24, 40, 64, 64),      # extents that are not tile multiples
    (0, 1, 64, 64, 32, 32),
    (0, 1, 32, 32, 32, 64),
    (0, 1, 32, 32, 64, 32),
    (0, 1, 32, 32, 128, 64),
    (0, 1, 32, 32, 64, 128),
    (0, 1, 16, 16, 256, 256),
    (0, 1, 16, 16, 128, 256),
    (0, 1, 16, 16, 256, 128),
    (0, 3, 8, 8, 128, 128),      # smaller than one tile
    (1, 2, 32, 32, 32, 32),
    (1, 1, 64, 64, 64, 64),
    (1, 2, 16, 16, 128, 128),
    (1, 1, 24, 40, 64, 64),
    (2, 2, 16, 16, 128, 128),
    (2, 1, 32, 32, 64, 64),
    (2, 1, 8, 24, 256, 256),
    (3, 2, 32, 32, 32, 64),
    (3, 1, 32, 32, 64, 128),
    (3, 1, 32, 32, 128, 64),
    (3, 2, 32, 32, 128, 128),
    (3, 1, 16, 16, 256, 256),
    (3, 2, 16, 24, 128, 384),    # fused q|k|v projection: three 128-column tiles
    (3, 1, 16, 16, 256, 768),
]


@pytest.mark.parametrize("mode,n,h,w,cin,cout", CONV_CASES)
def test_conv_umma(b200, mode, n, h, w, cin, cout):
    x = _rand_act(n, h, w, cin, 10 + mode)
    k = 1 if mode == 3 else 3
    wt, bias = _rand_conv(cout, cin, k, 20 + cin + cout)
    wp = b200.ops.pack_conv_weight(wt, 2 if mode == 2 else 0, DT)
    out = b200.ops.conv_umma(x, wp, bias, mode)
    assert out.dtype == DT
    ref = _ref_conv(x, wt, bias, mode)
    assert out.shape == ref.shape
    # mode 2 pre-sums bf16 weights (one extra rounding per weight) -> slightly looser
    _check_bf16(out, ref, f"conv mode {mode} {cin}->{cout}", rel=4e-3 if mode == 2 else 3e-3,
                ulp=2.0 ** -6 if mode == 2 else 2.0 ** -7)


@pytest.mark.parametrize("cin,cout,groups,out_f32,res_f32", [(128, 128, 16, True, True), (64, 64, 16, False, False),
                                                             (32, 32, 16, True, True), (64, 128, 32, True, False),
                                                             (256, 256, 32, False, True)])
def test_conv_umma_residual_and_stats(b200, cin, cout, groups, out_f32, res_f32):
    n, h, w = 2, 24, 24
    x = _rand_act(n, h, w, cin, 3)
    res = _rand_act(n, h, w, cout, 4)
    if res_f32:
        res = res.float() + 1e-3 * torch.randn(n, h, w, cout, device=DEV)
    wt, bias = _rand_conv(cout, cin, 3, 5)
    out, part = b200.ops.conv_umma(x, b200.ops.pack_conv_weight(wt, 0, DT), bias, 0, residual=res, gn_groups=groups,
                                   out_f32=out_f32)
    ref = _ref_conv(x, wt, bias, 0) + res.float()
    assert out.dtype == (torch.float32 if out_f32 else DT)
    if out_f32:  # fp32 accumulate + fp32 store: max-abs <= 1e-4 of the output scale
        assert float((out - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    else:
        _check_bf16(out, ref, "conv+residual")
    # statistics are those of the tensor that was written; partials summed in fixed order
    assert part.shape == (n, b200.ops.conv_parts(h, w, 0), groups, 2)
    acc = part.sum(dim=1)
    o = out.float().view(n, h * w, groups, cout // groups)
    s = o.sum(dim=(1, 3))
    q = (o * o).sum(dim=(1, 3))
    assert torch.allclose(acc[..., 0], s, rtol=1e-4, atol=1e-2), float((acc[..., 0] - s).abs().max())
    assert torch.allclose(acc[..., 1], q, rtol=1e-4, atol=1e-2), float((acc[..., 1] - q).abs().max())
    out2, part2 = b200.ops.conv_umma(x, b200.ops.pack_conv_weight(wt, 0, DT), bias, 0, residual=res, gn_groups=groups,
                                     out_f32=out_f32)
    assert torch.equal(out, out2) and torch.equal(part, part2), "conv + statistics must be run-to-run deterministic"


@pytest.mark.parametrize("mode,h,w,c", [(1, 32, 32, 64), (2, 16, 16, 128)])
def test_conv_umma_stats_strided_modes(b200, mode, h, w, c):
    x = _rand_act(2, h, w, c, 8)
    wt, bias = _rand_conv(c, c, 3, 9)
    out, part = b200.ops.conv_umma(x, b200.ops.pack_conv_weight(wt, 2 if mode == 2 else 0, DT), bias, mode, gn_groups=16,
                                   out_f32=True)
    assert part.shape[1] == b200.ops.conv_parts(h, w, mode)
    o = out.view(2, -1, 16, c // 16)
    assert torch.allclose(part.sum(1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(part.sum(1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("n,h,w,c,groups,emit16", [(2, 16, 16, 64, 32, True), (1, 32, 48, 64, 16, False), (2, 24, 40, 64, 32, True),
                                                   (2, 16, 16, 128, 32, True), (1, 40, 24, 128, 16, False), (3, 8, 8, 128, 32, True),
                                                   (1, 64, 64, 128, 32, True), (1, 20, 9, 64, 0, True)])
def test_up2x_conv3x3(b200, n, h, w, c, groups, emit16):
    """Halo-resident upsample+conv vs F.interpolate(nearest) + conv2d; statistics and the 16-bit copy of the stored values."""
    if DT != torch.float16:
        pytest.skip("fp16 operands only (bf16 takes conv_umma mode 2)")
    x = _rand_act(n, h, w, c, 21)
    wt, bias = _rand_conv(c, c, 3, 22)
    wp = b200.ops.pack_conv_weight(wt, 2, DT)
    r = b200.ops.up2x_conv3x3(x, wp, bias, gn_groups=groups, emit16=emit16)
    r = r if isinstance(r, tuple) else (r,)
    out = r[0]
    ref = _ref_conv(x, wt, bias, 2)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    _check_bf16(out, ref, f"up2x {c}", rel=4e-3, ulp=2.0 ** -6)      # pre-summed 16-bit weights: as conv_umma mode 2
    old = b200.ops.conv_umma(x, wp, bias, 2, out_f32=True)
    assert float((out - old).abs().max()) <= 1e-4 * max(1.0, float(old.abs().max())), "same operands, fp32 accumulate"
    if emit16:
        assert torch.equal(r[-1], out.to(DT))
    if groups:
        part = r[1]
        assert part.shape == (n, ((h + 15) // 16) * ((w + 15) // 16), groups, 2)
        o = out.view(n, -1, groups, c // groups)
        assert torch.allclose(part.sum(1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(part.sum(1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
        r2 = b200.ops.up2x_conv3x3(x, wp, bias, gn_groups=groups, emit16=emit16)
        assert torch.equal(r2[0], out) and torch.equal(r2[1], part), "deterministic"
    # 16-bit residual stream: only the fp16 tensor is written; statistics describe the stored (rounded) values
    r16 = b200.ops.up2x_conv3x3(x, wp, bias, gn_groups=groups, out_f32=False)
    o16 = r16[0] if groups else r16
    assert o16.dtype == DT and torch.equal(o16, out.to(DT))
    if groups:
        o = o16.float().view(n, -1, groups, c // groups)
        assert torch.allclose(r16[1].sum(1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(r16[1].sum(1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("n,h,w,c,groups,silu,f32", [(2, 32, 32, 128, 16, True, True), (2, 64, 64, 32, 16, True, False),
                                                     (1, 48, 16, 64, 16, False, True), (2, 16, 16, 256, 32, True, False),
                                                     (1, 8, 8, 64, 32, True, True), (3, 40, 24, 128, 32, True, True)])
def test_groupnorm(b200, n, h, w, c, groups, silu, f32):
    x = (_rand_act(n, h, w, c, 7).float() * 1.7 + 0.3)
    x = x + 1e-3 * torch.randn_like(x) if f32 else x.to(DT)
    gamma = torch.randn(c, device=DEV) * 0.5 + 1.0
    beta = torch.randn(c, device=DEV) * 0.2
    eps = 1e-6
    part = b200.ops.gn_stats(x, groups)
    assert torch.equal(part, b200.ops.gn_stats(x, groups)), "statistics must be deterministic"
    acc = part.sum(dim=1)
    xf = x.float().view(n, h * w, groups, c // groups)
    assert torch.allclose(acc[..., 0], xf.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(acc[..., 1], (xf * xf).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    ss = b200.ops.gn_finalize(part, gamma, beta, h * w, eps)
    y, raw = b200.ops.gn_apply(x, ss, silu, emit_raw=True, dtype=DT)
    assert torch.equal(raw, x.to(DT)) and y.dtype == DT
    assert torch.equal(y, b200.ops.gn_apply(x, ss, silu, dtype=DT))
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), groups, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 3, 1)
    _check_bf16(y, ref, "groupnorm(+silu)")
    # fp32-accumulate path: scale/shift against the fp32 statistics, max-abs <= 1e-4 (relative to scale ~1)
    mean = xf.mean(dim=(1, 3))
    var = xf.var(dim=(1, 3), unbiased=False)
    rstd = (var + eps).rsqrt()
    sc_ref = gamma.view(1, groups, -1) * rstd[..., None]
    sh_ref = beta.view(1, groups, -1) - mean[..., None] * sc_ref
    assert float((ss[..., 0].view(n, groups, -1) - sc_ref).abs().max()) <= 1e-4 * float(sc_ref.abs().max())
    assert float((ss[..., 1].view(n, groups, -1) - sh_ref).abs().max()) <= 1e-4 * max(1.0, float(sh_ref.abs().max()))


FUSED_CASES = [
    # (N, H, W, Cin, Cout, in_f32, norm, silu, res, out_f32, groups)
    (2, 32, 32, 128, 128, True, True, True, True, True, 16),
    (1, 16, 16, 128, 128, False, True, True, False, False, 16),
    (2, 40, 24, 64, 64, True, True, True, True, True, 16),     # extents not multiples of the tile
    (1, 64, 64, 32, 32, True, True, True, True, True, 16),
    (1, 32, 32, 32, 64, True, True, True, False, False, 16),
    (1, 32, 32, 64, 32, False, True, True, True, True, 16),
    (1, 32, 32, 128, 64, True, True, False, False, True, 16),
    (1, 32, 32, 64, 128, True, False, False, True, True, 32),
    (3, 8, 8, 128, 128, True, True, True, True, True, 16),      # image smaller than one tile
    (5, 48, 48, 64, 64, False, True, True, True, True, 16),     # > 1 tile per CTA not needed, many tiles
]


# shapes/modes that have a TMA-staged instantiation (conv_tma.cu): ResBlock conv1 (fp32 in, 16-bit out), conv2 to the
# fp32 stream (16-bit in, fp32 residual, fp32 out) and conv2 to a 16-bit operand; fp16 operands, Cin/Cout <= 64
TMA_CASES = [
    # (N, H, W, Cin, Cout, in_f32, res, out_f32, groups)
    (2, 32, 32, 32, 32, True, False, False, 16),
    (2, 32, 32, 32, 32, False, True, True, 16),
    (1, 48, 40, 32, 32, False, True, False, 16),     # edge tiles (extents not multiples of 16)
    (2, 24, 24, 64, 64, True, False, False, 16),
    (3, 40, 24, 64, 64, False, True, True, 32),
    (1, 32, 32, 32, 64, True, False, False, 16),
    (1, 32, 32, 64, 32, True, False, False, 16),
    (2, 16, 16, 64, 32, False, True, True, 16),
    (1, 8, 8, 32, 64, False, True, False, 32),       # image smaller than a tile
    (9, 64, 64, 32, 32, False, True, True, 16),      # several tiles per CTA on any SM count >= 16
    # 16-bit residual stream: conv1 (16-bit in, 16-bit out) and conv2 (16-bit residual added in place, 16-bit out)
    (2, 32, 32, 32, 32, False, False, False, 16),
    (2, 32, 32, 32, 32, False, "h16", False, 16),
    (1, 48, 40, 32, 32, False, "h16", False, 16),
    (3, 40, 24, 64, 64, False, "h16", False, 32),
    (1, 24, 24, 64, 32, False, False, False, 16),
    (1, 8, 8, 32, 64, False, "h16", False, 32),
    (9, 64, 64, 32, 32, False, "h16", False, 16),
]


@pytest.mark.parametrize("n,h,w,cin,cout,in_f32,res,out_f32,groups", TMA_CASES)
def test_conv3x3_fused_tma(b200, n, h, w, cin, cout, in_f32, res, out_f32, groups):
    if DT != torch.float16:
        pytest.skip("the TMA-staged kernel is instantiated for fp16 operands only")
    x = _rand_act(n, h, w, cin, 41).float() * 1.5 + 0.2
    x = x + 1e-3 * torch.randn_like(x) if in_f32 else x.to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 42)
    ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    xin = F.silu(x.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    r = torch.randn(n, h, w, cout, device=DEV) if res else None
    if res == "h16":
        r = r.to(DT)
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    wp = b200.ops.pack_conv_weight(wt, 0, DT)
    outs = {}
    try:
        for impl in (2, 1):
            b200.ops.FUSED_IMPL = impl
            out, part = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
            out2, part2 = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
            assert torch.equal(out, out2) and torch.equal(part, part2), f"impl {impl} not deterministic"
            _check_bf16(out.to(DT), ref, f"fused conv impl={impl}")
            o = out.float().view(n, h * w, groups, cout // groups)
            acc = part.sum(dim=1)
            assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            outs[impl] = out.float()
    finally:
        b200.ops.FUSED_IMPL = 0
    # both implementations compute the same thing (fp32 add order differs by at most an ulp or a 16-bit flip)
    assert float((outs[1] - outs[2]).abs().max()) <= 2.0 ** -9 * float(ref.abs().max())


TMA2_SHAPES = [(32, 32), (32, 64), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128)]
# (in_f32, res, out_f32): conv1 | conv2 | conv2 -> operand | conv1, conv2 on a 16-bit residual stream
TMA2_MODES = [(True, False, False), (False, True, True), (False, True, False), (False, False, False), (False, "h16", False)]


@pytest.mark.parametrize("cin,cout", TMA2_SHAPES)
@pytest.mark.parametrize("in_f32,res,out_f32", TMA2_MODES)
@pytest.mark.parametrize("n,h,w,groups", [(2, 32, 32, 32), (1, 40, 24, 16), (5, 48, 48, 32)])
def test_conv3x3_fused_tma2(b200, cin, cout, in_f32, res, out_f32, n, h, w, groups):
    """Chunk-pipelined TMA kernel (impl 3) vs the fp32 reference and vs the register-staged kernel (impl 1)."""
    if DT != torch.float16:
        pytest.skip("the TMA-staged kernels are instantiated for fp16 operands only")
    groups = min(groups, cout // 2)
    x = _rand_act(n, h, w, cin, 51).float() * 1.5 + 0.2
    x = x + 1e-3 * torch.randn_like(x) if in_f32 else x.to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 52)
    ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    xin = F.silu(x.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    r = torch.randn(n, h, w, cout, device=DEV) if res else None
    if res == "h16":
        r = r.to(DT)
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    wp = b200.ops.pack_conv_weight(wt, 0, DT)
    outs = {}
    try:
        for impl in (3, 1):
            b200.ops.FUSED_IMPL = impl
            out, part = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
            out2, part2 = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
            assert torch.equal(out, out2) and torch.equal(part, part2), f"impl {impl} not deterministic"
            # the 16-bit stream modes of impl 3 run the prologue in packed half2 (h + h*tanh(h), three fp16 roundings): twice
            # the tolerances, as for the row-band kernel
            h2 = impl == 3 and not in_f32 and not out_f32 and res in (False, "h16")
            _check_bf16(out.to(DT), ref, f"fused conv impl={impl}", rel=6e-3 if h2 else 3e-3, ulp=2.0 ** -6 if h2 else 2.0 ** -7)
            o = out.float().view(n, h * w, groups, cout // groups)
            acc = part.sum(dim=1)
            assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            outs[impl] = out.float()
    finally:
        b200.ops.FUSED_IMPL = 0
    stream16 = not in_f32 and not out_f32 and res in (False, "h16")
    assert float((outs[1] - outs[3]).abs().max()) <= 2.0 ** (-8 if stream16 else -9) * float(ref.abs().max())


@pytest.mark.parametrize("cin,cout", [(128, 256), (256, 128), (256, 256)])
@pytest.mark.parametrize("in_f32,res,out_f32", TMA2_MODES)
@pytest.mark.parametrize("n,h,w,groups", [(2, 32, 32, 32), (1, 40, 24, 16), (5, 48, 48, 32), (3, 8, 8, 32)])
def test_conv3x3_fused_tma2_wide(b200, cin, cout, in_f32, res, out_f32, n, h, w, groups):
    """256-wide layers (config B) on the chunk-pipelined kernel (one accumulator stage for 256 outputs, 2-chunk ring) vs the
    fp32 reference and vs the unfused route (gn_apply + conv_umma).  The 16-bit stream forms always exist; the fp32-stream
    forms only where their staging fits in shared memory (ptivae_conv3x3_fused_query), and the rest raise."""
    if DT != torch.float16:
        pytest.skip("the TMA-staged kernels are instantiated for fp16 operands only")
    ops = b200.ops
    rdt = None if not res else (DT if res == "h16" else torch.float32)
    ok = ops.conv3x3_fused_supported(torch.float32 if in_f32 else DT, rdt, out_f32, cin, cout, DT)
    if not in_f32 and not out_f32 and rdt != torch.float32:
        assert ok, "the 16-bit stream forms must be instantiated"
    x = _rand_act(n, h, w, cin, 61).float() * 1.5 + 0.2
    x = x + 1e-3 * torch.randn_like(x) if in_f32 else x.to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 62)
    ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    r = torch.randn(n, h, w, cout, device=DEV) if res else None
    if res == "h16":
        r = r.to(DT)
    wp = ops.pack_conv_weight(wt, 0, DT)
    if not ok:
        with pytest.raises(b200._lib.PtivaeError):
            ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
        return
    xin = F.silu(x.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    out, part = ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
    out2, part2 = ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=out_f32)
    assert torch.equal(out, out2) and torch.equal(part, part2), "not deterministic"
    h2 = not in_f32 and not out_f32 and res in (False, "h16")    # packed-half2 prologue on the 16-bit stream
    _check_bf16(out.to(DT), ref, "fused conv 256", rel=6e-3 if h2 else 3e-3, ulp=2.0 ** -6 if h2 else 2.0 ** -7)
    o = out.float().view(n, h * w, groups, cout // groups)
    acc = part.sum(dim=1)
    assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    y = ops.gn_apply(x, ss, silu=True, dtype=DT)
    un = ops.conv_umma(y, wp, bias, 0, residual=r, out_f32=out_f32)
    assert float((un.float() - out.float()).abs().max()) <= 2.0 ** (-7 if h2 else -8) * float(ref.abs().max())


@pytest.mark.parametrize("cin,cout", [(128, 128), (128, 256), (256, 128), (256, 256)])
@pytest.mark.parametrize("res", [False, "h16"])
@pytest.mark.parametrize("n,h,w,groups", [(2, 32, 32, 32), (1, 40, 24, 16), (5, 48, 48, 32), (3, 8, 8, 32), (40, 32, 32, 16), (1, 16, 16, 32)])
def test_conv3x3_fused_pair(b200, cin, cout, res, n, h, w, groups):
    """Two-SM kernel (impl 5: cta_group::2 MMAs, a pair of CTAs per pair of tiles) vs the fp32 reference and vs the
    one-SM chunk-pipelined kernel (impl 3, same MMA order): odd tile counts (ghost tile), several iterations per pair,
    a single tile, ragged edges."""
    if DT != torch.float16:
        pytest.skip("the TMA-staged kernels are instantiated for fp16 operands only")
    ops = b200.ops
    x = (_rand_act(n, h, w, cin, 81).float() * 1.5 + 0.2).to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 82)
    ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    xin = F.silu(x.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    r = torch.randn(n, h, w, cout, device=DEV).to(DT) if res else None
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    wp = ops.pack_conv_weight(wt, 0, DT)
    outs = {}
    try:
        for impl in (5, 3):
            ops.FUSED_IMPL = impl
            out, part = ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=False)
            out2, part2 = ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=False)
            assert torch.equal(out, out2) and torch.equal(part, part2), f"impl {impl} not deterministic"
            _check_bf16(out, ref, f"fused conv impl={impl}", rel=6e-3, ulp=2.0 ** -6)   # packed-half2 prologue in both
            o = out.float().view(n, h * w, groups, cout // groups)
            acc = part.sum(dim=1)
            assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            outs[impl] = (out, part)
    finally:
        ops.FUSED_IMPL = 0
    assert float((outs[5][0].float() - outs[3][0].float()).abs().max()) <= 2.0 ** -9 * float(ref.abs().max())


@pytest.mark.parametrize("cin", [32, 64])
@pytest.mark.parametrize("res", [False, "h16"])
@pytest.mark.parametrize("n,h,w,groups", [(2, 16, 128, 16), (1, 40, 200, 16), (3, 64, 256, 8), (1, 7, 130, 16), (200, 16, 128, 16),
                                          (1, 260, 256, 16)])
def test_conv3x3_fused_band(b200, cin, res, n, h, w, groups):
    """Row-band kernel (impl 4: 32 output channels, 16-bit in/out, rows streamed through a ring) vs the fp32 reference and
    vs the register-staged kernel (impl 1): multi-segment CTAs, ring wrap-around, ragged right / bottom edges."""
    if DT != torch.float16:
        pytest.skip("the row-band kernel is instantiated for fp16 operands only")
    cout = 32
    x = (_rand_act(n, h, w, cin, 71).float() * 1.5 + 0.2).to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 72)
    ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    xin = F.silu(x.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    r = torch.randn(n, h, w, cout, device=DEV).to(DT) if res else None
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    wp = b200.ops.pack_conv_weight(wt, 0, DT)
    outs = {}
    try:
        for impl in (4, 1):
            b200.ops.FUSED_IMPL = impl
            out, part = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=False)
            out2, part2 = b200.ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=groups, out_f32=False)
            assert torch.equal(out, out2) and torch.equal(part, part2), f"impl {impl} not deterministic"
            # impl 4 evaluates SiLU as h + h*tanh(h) with tanh.approx.f32 (relative error 2^-11): twice the tolerances
            _check_bf16(out, ref, f"fused conv impl={impl}", rel=6e-3 if impl == 4 else 3e-3, ulp=2.0 ** -6 if impl == 4 else 2.0 ** -7)
            o = out.float().view(n, h * w, groups, cout // groups)
            acc = part.sum(dim=1)
            assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2), impl
            outs[impl] = out.float()
        print(f"band vs fp32 reference rel-L2 {_rel_l2(outs[4], ref):.2e} (register-staged kernel: {_rel_l2(outs[1], ref):.2e})")
        b200.ops.FUSED_IMPL = 4     # no statistics, identity prologue
        o_plain = b200.ops.conv3x3_fused(x, None, False, wp, bias, residual=r, out_f32=False)
        ref_plain = F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=1).permute(0, 2, 3, 1)
        _check_bf16(o_plain, ref_plain + (r.float() if res else 0.0), "band, identity prologue")
    finally:
        b200.ops.FUSED_IMPL = 0
    assert float((outs[1] - outs[4]).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())


@pytest.mark.parametrize("c,sc", [(32, 64), (64, 32)])
@pytest.mark.parametrize("n,h,w,groups", [(2, 32, 32, 16), (1, 40, 24, 16), (5, 48, 48, 32)])
def test_conv3x3_fused_shortcut(b200, c, sc, n, h, w, groups):
    """conv2 + fused 1x1 shortcut vs conv2d + conv2d in fp32, and vs the two-kernel route (shortcut conv -> residual)."""
    if DT != torch.float16:
        pytest.skip("fp16 operands only")
    groups = min(groups, c // 2)
    hh = (_rand_act(n, h, w, c, 61).float() * 1.5 + 0.2).to(DT)
    xr = _rand_act(n, h, w, sc, 62)
    wt, bias = _rand_conv(c, c, 3, 63)
    wsc, bsc = _rand_conv(c, sc, 1, 64)
    ss = (torch.randn(n, c, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
    xin = F.silu(hh.float().permute(0, 3, 1, 2) * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]).to(DT).float()
    ref = (F.conv2d(xin, wt, bias, padding=1) + F.conv2d(xr.float().permute(0, 3, 1, 2), wsc, bsc)).permute(0, 2, 3, 1)
    assert b200.ops.fused_sc_supported(hh.dtype, c, c, sc)
    wp, wscp = b200.ops.pack_conv_weight(wt, 0, DT), b200.ops.pack_conv_weight(wsc, 0, DT)
    out, part = b200.ops.conv3x3_fused_sc(hh, ss, True, wp, bias + bsc, xr, wscp, gn_groups=groups)
    out2, part2 = b200.ops.conv3x3_fused_sc(hh, ss, True, wp, bias + bsc, xr, wscp, gn_groups=groups)
    assert torch.equal(out, out2) and torch.equal(part, part2)
    _check_bf16(out.to(DT), ref, "fused conv + shortcut")
    o = out.view(n, h * w, groups, c // groups)
    acc = part.sum(dim=1)
    assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    scv = b200.ops.conv_umma(xr, wscp, bsc, 3, out_f32=True)
    two, _ = b200.ops.conv3x3_fused(hh, ss, True, wp, bias, residual=scv, gn_groups=groups, out_f32=True)
    assert float((out - two).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    # 16-bit residual stream: rounded once on the way out, statistics of the stored values; this form runs the prologue in
    # packed half2 (h + h*tanh(h)) like every 16-bit-stream mode: twice the tolerances against the fp32 reference
    o16, p16 = b200.ops.conv3x3_fused_sc(hh, ss, True, wp, bias + bsc, xr, wscp, gn_groups=groups, out_f32=False)
    o16b, p16b = b200.ops.conv3x3_fused_sc(hh, ss, True, wp, bias + bsc, xr, wscp, gn_groups=groups, out_f32=False)
    assert o16.dtype == DT and torch.equal(o16, o16b) and torch.equal(p16, p16b)
    _check_bf16(o16, ref, "fused conv + shortcut, 16-bit out", rel=6e-3, ulp=2.0 ** -6)
    assert float((o16.float() - out).abs().max()) <= 2.0 ** -7 * float(ref.abs().max())
    o = o16.float().view(n, h * w, groups, c // groups)
    assert torch.allclose(p16.sum(dim=1)[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(p16.sum(dim=1)[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("n,h,w,cin,cout,in_f32,norm,silu,res,out_f32,groups", FUSED_CASES)
def test_conv3x3_fused(b200, n, h, w, cin, cout, in_f32, norm, silu, res, out_f32, groups):
    x = _rand_act(n, h, w, cin, 31).float() * 1.5 + 0.2
    x = x + 1e-3 * torch.randn_like(x) if in_f32 else x.to(DT)
    wt, bias = _rand_conv(cout, cin, 3, 32)
    ss = None
    xin = x.float().permute(0, 3, 1, 2)
    if norm:
        ss = (torch.randn(n, cin, 2, device=DEV) * 0.5 + torch.tensor([1.0, 0.0], device=DEV)).contiguous()
        xin = xin * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]
        if silu:
            xin = F.silu(xin)
    xin = xin.to(DT).float()          # the operand the tensor cores see
    r = None
    if res:
        r = torch.randn(n, h, w, cout, device=DEV)
    ref = F.conv2d(xin, wt, bias, padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r
    out, part = b200.ops.conv3x3_fused(x, ss, silu, b200.ops.pack_conv_weight(wt, 0, DT), bias, residual=r,
                                       gn_groups=groups, out_f32=out_f32)
    if out_f32:
        # operand rounding can flip (silu differs in the last ulp): allow 1 ulp of the operand format on the sum
        _check_bf16(out.to(DT), ref, "fused conv (fp32 out)")
    else:
        # 16-bit in and out without an fp32 residual = a 16-bit-stream mode: packed-half2 prologue (h + h*tanh(h)), twice
        # the tolerances
        h2 = not in_f32 and not res and norm
        _check_bf16(out, ref, "fused conv", rel=6e-3 if h2 else 3e-3, ulp=2.0 ** -6 if h2 else 2.0 ** -7)
    o = out.float().view(n, h * w, groups, cout // groups)
    acc = part.sum(dim=1)
    assert torch.allclose(acc[..., 0], o.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(acc[..., 1], (o * o).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
    out2, part2 = b200.ops.conv3x3_fused(x, ss, silu, b200.ops.pack_conv_weight(wt, 0, DT), bias, residual=r,
                                         gn_groups=groups, out_f32=out_f32)
    assert torch.equal(out, out2) and torch.equal(part, part2)


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 1, 32, 64, 64), (1, 1, 64, 32, 48), (2, 4, 128, 16, 16), (1, 10, 256, 8, 8),
                                           (1, 1, 32, 37, 50)])
def test_conv_small_cin(b200, n, cin, cout, h, w):
    x = torch.randn(n, cin, h, w, device=DEV)
    wt, bias = _rand_conv(cout, cin, 3, 9)
    out = b200.ops.conv3x3_small_cin(x, wt, bias, dtype=DT)
    ref = F.conv2d(x, wt, bias, padding=1).permute(0, 2, 3, 1)
    _check_bf16(out, ref, "small_cin")
    out32 = b200.ops.conv3x3_small_cin(x, wt, bias)
    assert float((out32 - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))
    for groups in (32, 16, 8):
        if not b200.ops.small_cin_stats_supported(cin, cout, groups):
            continue
        for dt in (torch.float32, DT):
            o, part = b200.ops.conv3x3_small_cin(x, wt, bias, dtype=dt, gn_groups=groups)
            assert torch.equal(o, b200.ops.conv3x3_small_cin(x, wt, bias, dtype=dt))
            of = o.float().view(n, h * w, groups, cout // groups)
            acc = part.sum(dim=1)
            assert torch.allclose(acc[..., 0], of.sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
            assert torch.allclose(acc[..., 1], (of * of).sum(dim=(1, 3)), rtol=1e-4, atol=1e-2)
            assert torch.equal(part, b200.ops.conv3x3_small_cin(x, wt, bias, dtype=dt, gn_groups=groups)[1])


@pytest.mark.parametrize("n,cin,cout,h,w,norm,f32", [(2, 32, 1, 64, 64, True, True), (1, 64, 1, 32, 32, True, False),
                                                     (2, 128, 4, 16, 16, True, True), (1, 256, 10, 8, 8, True, True),
                                                     (1, 32, 1, 16, 16, False, False), (1, 32, 1, 40, 70, True, False),
                                                     (2, 64, 2, 100, 36, True, True)])
def test_conv_small_cout(b200, n, cin, cout, h, w, norm, f32):
    x = _rand_act(n, h, w, cin, 11)
    if f32:
        x = x.float() + 1e-3 * torch.randn(n, h, w, cin, device=DEV)
    wt, bias = _rand_conv(cout, cin, 3, 12)
    ss = None
    xin = x.float().permute(0, 3, 1, 2)
    if norm:
        ss = torch.randn(n, cin, 2, device=DEV) * 0.5 + 0.5
        xin = xin * ss[:, :, 0, None, None] + ss[:, :, 1, None, None]
    out = b200.ops.conv3x3_small_cout(x, wt, bias, ss)
    ref = F.conv2d(xin, wt, bias, padding=1)
    # fp32 accumulate, fp32 out: max-abs <= 1e-4 of the output scale
    assert float((out - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max()))


def test_conv1x1_small_and_sigma(b200):
    x = torch.randn(3, 4, 16, 16, device=DEV) * 8
    w = torch.randn(4, 4, 1, 1, device=DEV)
    b = torch.randn(4, device=DEV)
    mu = b200.ops.conv1x1_small(x, w, b, 0)
    assert float((mu - F.conv2d(x, w, b)).abs().max()) <= 1e-4 * 30
    sg = b200.ops.conv1x1_small(x, w * 3, b, 1)
    ref = torch.exp(torch.clamp(F.conv2d(x, w * 3, b), -30.0, 20.0) / 2)
    assert torch.allclose(sg, ref, rtol=2e-5, atol=1e-12)
    assert float(sg.max()) <= math.exp(10.0) * 1.0001 and float(sg.min()) >= math.exp(-15.0) * 0.9999


@pytest.mark.parametrize("b,l,d", [(2, 1024, 128), (1, 256, 128), (2, 320, 128), (1, 128, 128), (1, 512, 256),
                                   (1, 200, 256), (2, 256, 64)])
def test_attention(b200, b, l, d):
    g = torch.Generator().manual_seed(l + d)
    q, k, v = [(torch.randn(b, l, d, generator=g) * s).to(DEV).to(DT) for s in (1.0, 1.0, 1.0)]
    out = b200.ops.attention(q, k, v)
    att = torch.softmax(torch.einsum("bxd,byd->bxy", q.float(), k.float()) * d ** -0.5, dim=-1)
    ref = torch.einsum("bxy,byd->bxd", att, v.float())
    # P is rounded to 16 bit before PV: allow 2 ulps of the output scale
    _check_bf16(out, ref, f"attention L={l} d={d}", rel=6e-3, ulp=2.0 ** -6)


def test_attention_reads_fused_qkv_slices(b200):
    """q, k, v as channel slices of one [B, L, 3D] projection (row stride 3D) == separate contiguous tensors."""
    b, l, d = 2, 320, 128
    qkv = (torch.randn(b, l, 3 * d, device=DEV) * 0.7).to(DT)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    fused = b200.ops.attention(q, k, v)
    sep = b200.ops.attention(q.contiguous(), k.contiguous(), v.contiguous())
    assert torch.equal(fused, sep)
    with pytest.raises(b200._lib.PtivaeError):
        b200.ops.attention(q, k.contiguous(), v)


def test_attention_peaked(b200):
    # large-magnitude scores exercise the online-softmax rescaling across key blocks
    b, l, d = 1, 512, 128
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(b, l, d, generator=g) * 4).to(DEV).to(DT)
    k = (torch.randn(b, l, d, generator=g) * 4).to(DEV).to(DT)
    v = torch.randn(b, l, d, generator=g).to(DEV).to(DT)
    out = b200.ops.attention(q, k, v)
    att = torch.softmax(torch.einsum("bxd,byd->bxy", q.float(), k.float()) * d ** -0.5, dim=-1)
    ref = torch.einsum("bxy,byd->bxd", att, v.float())
    _check_bf16(out, ref, "attention peaked", rel=8e-3, ulp=2.0 ** -5)


def test_latent_sample(b200, oracle):
    mu = torch.randn(3, 4, 8, 8, device=DEV)
    sg = torch.rand(3, 4, 8, 8, device=DEV) + 0.1
    eps = torch.randn(3, 4, 8, 8, device=DEV)
    z = b200.ops.latent_sample(mu, sg, eps=eps)
    assert torch.allclose(z, mu + eps * sg, rtol=1e-6, atol=1e-6)
    # the product's own generator, pinned against the numpy restatement of Philox4x32-10 + Box-Muller
    n = mu.numel()
    z2, e2 = b200.ops.latent_sample(mu, sg, seed=1234, offset=7, return_eps=True)
    ref = torch.from_numpy(oracle.philox_normal(n, 1234, 7)).to(DEV).view_as(mu)
    assert float((e2 - ref).abs().max()) <= 2e-4
    assert torch.allclose(z2, mu + e2 * sg, rtol=1e-6, atol=1e-6)
    # device-resident state gives the same stream and advances
    st = torch.tensor([1234, 7], device=DEV, dtype=torch.int64)
    z3, e3 = b200.ops.latent_sample(mu, sg, rng_dev=st, return_eps=True)
    assert torch.equal(e3, e2)
    b200.ops.rng_advance(st)
    assert st.tolist() == [1234, 8]
    big = torch.zeros(1 << 20, device=DEV)
    _, e4 = b200.ops.latent_sample(big, big + 1, seed=99, offset=0, return_eps=True)
    assert abs(float(e4.mean())) < 5e-3 and abs(float(e4.std()) - 1.0) < 5e-3


def test_losses(b200, oracle):
    g = torch.Generator().manual_seed(0)
    mu = torch.randn(5, 4, 32, 32, generator=g)
    sg = torch.rand(5, 4, 32, 32, generator=g) * 2 + 0.01
    for flag in (True, False):
        ref = oracle.kl_loss_ref(mu, sg, input_is_logvar=flag)
        got = b200.compute_kl_loss(mu.to(DEV), sg.to(DEV), input_is_logvar=flag)
        assert abs(float(got) - float(ref)) <= 1e-3 * abs(float(ref)), (flag, float(got), float(ref))
    a = torch.randn(3, 1, 64, 64, generator=g)
    b = torch.randn(3, 1, 64, 64, generator=g)
    assert abs(float(b200.l1_loss(a.to(DEV), b.to(DEV))) - float(oracle.l1_ref(a, b))) <= 1e-3 * float(oracle.l1_ref(a, b))
    assert abs(float(b200.mse_loss(a.to(DEV), b.to(DEV))) - float(oracle.l2_ref(a, b))) <= 1e-3 * float(oracle.l2_ref(a, b))
    # determinism: bit-identical across repeated launches
    r1 = b200.ops.l1l2(a.to(DEV), b.to(DEV))
    r2 = b200.ops.l1l2(a.to(DEV), b.to(DEV))
    assert torch.equal(r1, r2)
    # reference overflow behaviour is preserved: exp(sigma) with huge sigma -> inf
    hot = torch.full((1, 4, 4, 4), 200.0)
    assert math.isinf(float(b200.compute_kl_loss(torch.zeros_like(hot).to(DEV), hot.to(DEV))))
    assert math.isinf(float(oracle.kl_loss_ref(torch.zeros_like(hot), hot)))
