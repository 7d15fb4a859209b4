"""Event-timed micro-benchmarks of single ops at the bench shapes: bench_ops.py [name ...]"""
import sys, pathlib, math
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
b200 = _pkg.load(); ops = b200.ops
N = 64
DT = torch.float16


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cout1():
    x = torch.randn(N, 256, 256, 32, device="cuda").to(DT); w = torch.randn(1, 32, 3, 3, device="cuda"); b = torch.randn(1, device="cuda")
    ss = torch.randn(N, 32, 2, device="cuda")
    return timeit(lambda: ops.conv3x3_small_cout(x, w, b, ss)), x.numel() * 2 + N * 65536 * 4


def cout4():
    x = torch.randn(N, 32, 32, 128, device="cuda"); w = torch.randn(4, 128, 3, 3, device="cuda"); b = torch.randn(4, device="cuda")
    ss = torch.randn(N, 128, 2, device="cuda")
    return timeit(lambda: ops.conv3x3_small_cout(x, w, b, ss)), x.numel() * 4


def cin1():
    x = torch.randn(N, 1, 256, 256, device="cuda"); w = torch.randn(32, 1, 3, 3, device="cuda"); b = torch.randn(32, device="cuda")
    return timeit(lambda: ops.conv3x3_small_cin(x, w, b)), N * 65536 * 32 * 4


def cin4():
    x = torch.randn(N, 4, 32, 32, device="cuda"); w = torch.randn(128, 4, 3, 3, device="cuda"); b = torch.randn(128, device="cuda")
    return timeit(lambda: ops.conv3x3_small_cin(x, w, b)), N * 1024 * 128 * 4


def up(c, hw, emit16):
    x = torch.randn(N, hw, hw, c, device="cuda").to(DT)
    w = torch.randn(c, c, 3, 3, device="cuda") / math.sqrt(9 * c); b = torch.randn(c, device="cuda")
    wp = ops.pack_conv_weight(w, 2, DT)
    by = x.numel() * 2 + N * 4 * hw * hw * c * (4 + (2 if emit16 else 0))
    return timeit(lambda: ops.conv_umma(x, wp, b, 2, gn_groups=32, out_f32=True, emit16=emit16)), by


def up_new(c, hw, emit16):
    x = torch.randn(N, hw, hw, c, device="cuda").to(DT)
    w = torch.randn(c, c, 3, 3, device="cuda") / math.sqrt(9 * c); b = torch.randn(c, device="cuda")
    wp = ops.pack_conv_weight(w, 2, DT)
    by = x.numel() * 2 + N * 4 * hw * hw * c * (4 + (2 if emit16 else 0))
    return timeit(lambda: ops.up2x_conv3x3(x, wp, b, gn_groups=32, emit16=emit16)), by


def fused(cin, cout, hw, conv2, out32=True, impl=0):
    ops.FUSED_IMPL = impl
    x = torch.randn(N, hw, hw, cin, device="cuda", dtype=DT if conv2 else torch.float32)
    wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin), 0, DT)
    bias = torch.randn(cout, device="cuda"); ss = torch.randn(N, cin, 2, device="cuda")
    res = torch.randn(N, hw, hw, cout, device="cuda") if conv2 else None
    o32 = bool(conv2 and out32)
    by = x.numel() * x.element_size() + N * hw * hw * cout * ((4 if o32 else 2) + (4 if conv2 else 0))
    return timeit(lambda: ops.conv3x3_fused(x, ss, True, wp, bias, residual=res, gn_groups=min(32, cout // 2), out_f32=o32)), by


def fused16(cin, cout, hw, res, impl=0):
    """ResBlock conv on the fp16 stream: 16-bit in, 16-bit out, optional 16-bit residual."""
    ops.FUSED_IMPL = impl
    x = torch.randn(N, hw, hw, cin, device="cuda").to(DT)
    wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin), 0, DT)
    bias = torch.randn(cout, device="cuda"); ss = torch.randn(N, cin, 2, device="cuda")
    r = torch.randn(N, hw, hw, cout, device="cuda").to(DT) if res else None
    by = x.numel() * 2 + N * hw * hw * cout * 2 * (2 if res else 1)
    ms = timeit(lambda: ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=min(32, cout // 2), out_f32=False))
    ops.FUSED_IMPL = 0
    print(f"   {cin}->{cout}@{hw} res={int(res)}: {2.0 * N * hw * hw * cin * cout * 9 / ms / 1e9:.0f} TFLOP/s")
    return ms, by


def gnbwd(n, hw, c, res=True):
    """GroupNorm+SiLU backward as the training step calls it: fp32 x, bf16 dA, fp32 residual gradient -> dx fp32 + bf16,
    re-materialised activation, fused bias gradient.  bytes = reads (x twice, dA twice, residual) + writes."""
    x = torch.randn(n, hw, hw, c, device="cuda"); da = torch.randn(n, hw, hw, c, device="cuda").to(torch.bfloat16)
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    r = torch.randn(n, hw, hw, c, device="cuda") if res else None
    ss, mr = ops.gn_finalize(ops.gn_stats(x, 16), gamma, beta, hw * hw, 1e-6, return_mean_rstd=True)
    dg, db, cs = torch.empty(c, device="cuda"), torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    e = n * hw * hw * c
    by = e * (2 * 4 + 2 * 2 + (4 if res else 0) + 4 + 2 + 2)
    return timeit(lambda: ops.gn_bwd(x, da, ss, mr, gamma, True, dg, db, residual=r, want_act=True, colsum_out=cs)), by


def fused_sc(c, sc, hw):
    h = torch.randn(N, hw, hw, c, device="cuda").to(DT); xr = torch.randn(N, hw, hw, sc, device="cuda").to(DT)
    wp = ops.pack_conv_weight(torch.randn(c, c, 3, 3, device="cuda") / math.sqrt(9 * c), 0, DT)
    wsc = ops.pack_conv_weight(torch.randn(c, sc, 1, 1, device="cuda") / math.sqrt(sc), 0, DT)
    bias = torch.randn(c, device="cuda"); ss = torch.randn(N, c, 2, device="cuda")
    by = h.numel() * 2 + xr.numel() * 2 + N * hw * hw * c * 4
    return timeit(lambda: ops.conv3x3_fused_sc(h, ss, True, wp, bias, xr, wsc, gn_groups=min(32, c // 2))), by


def attn(l, d, b=64):
    q, k, v = (torch.randn(b, l, d, device="cuda").to(DT) for _ in range(3))
    ms, _ = timeit(lambda: ops.attention(q, k, v)), 0
    print(f"   attention L={l} d={d}: {4.0 * b * l * l * d / ms / 1e9:.0f} TFLOP/s")
    return ms, 8 * b * l * d


def band(cin, res):
    """row-band kernel (impl 4): 16-bit in / out, optional 16-bit residual; bytes = SURVEY 8(d) count (+ residual)."""
    ops.FUSED_IMPL = 4
    x = torch.randn(N, 256, 256, cin, device="cuda").to(DT)
    wp = ops.pack_conv_weight(torch.randn(32, cin, 3, 3, device="cuda") / math.sqrt(9 * cin), 0, DT)
    bias = torch.randn(32, device="cuda"); ss = torch.randn(N, cin, 2, device="cuda")
    r = torch.randn(N, 256, 256, 32, device="cuda").to(DT) if res else None
    by = x.numel() * 2 + N * 65536 * 32 * 2 * (2 if res else 1)
    ms = timeit(lambda: ops.conv3x3_fused(x, ss, True, wp, bias, residual=r, gn_groups=16, out_f32=False))
    ops.FUSED_IMPL = 0
    print(f"   band {cin}->32: {2.0 * N * 65536 * cin * 32 * 9 / ms / 1e9:.0f} TFLOP/s")
    return ms, by


def gnstats(c, hw, dt):
    x = torch.randn(N, hw, hw, c, device="cuda").to(dt)
    return timeit(lambda: ops.gn_stats(x, 16)), x.numel() * x.element_size()


def gnapply(c, hw, dt):
    x = torch.randn(N, hw, hw, c, device="cuda").to(dt); ss = torch.randn(N, c, 2, device="cuda")
    return timeit(lambda: ops.gn_apply(x, ss, silu=True, dtype=DT)), x.numel() * (x.element_size() + 2)


def l1l2():
    a, b = torch.randn(N, 1, 256, 256, device="cuda"), torch.randn(N, 1, 256, 256, device="cuda")
    return timeit(lambda: ops.l1l2(a, b)), a.numel() * 8


def kl():
    mu, sg = torch.randn(N, 4, 32, 32, device="cuda"), torch.rand(N, 4, 32, 32, device="cuda") + 0.5
    return timeit(lambda: ops.kl_loss(mu, sg, True)), mu.numel() * 8


def sample():
    mu, sg = torch.randn(N, 4, 32, 32, device="cuda"), torch.rand(N, 4, 32, 32, device="cuda") + 0.5
    return timeit(lambda: ops.latent_sample(mu, sg, seed=1, offset=1)), mu.numel() * 12


CASES = {
    "s64c1": lambda: fused16(64, 64, 128, False), "s64c2": lambda: fused16(64, 64, 128, True), "s12864": lambda: fused16(128, 64, 128, False),
    "s128c1": lambda: fused16(128, 128, 64, False), "s128c2": lambda: fused16(128, 128, 64, True),
    "s128c1s": lambda: fused16(128, 128, 32, False), "s128c2s": lambda: fused16(128, 128, 32, True), "s3264": lambda: fused16(32, 64, 128, False),
    "s256c1": lambda: fused16(256, 256, 64, False), "s256c2": lambda: fused16(256, 256, 64, True),
    "s128256": lambda: fused16(128, 256, 64, False), "s256128": lambda: fused16(256, 128, 128, False),
    "p128c1": lambda: fused16(128, 128, 64, False, impl=5), "p128c2": lambda: fused16(128, 128, 64, True, impl=5),
    "p128c1s": lambda: fused16(128, 128, 32, False, impl=5), "p128c2s": lambda: fused16(128, 128, 32, True, impl=5),
    "p256c1": lambda: fused16(256, 256, 64, False, impl=5), "p256c2": lambda: fused16(256, 256, 64, True, impl=5),
    "p128256": lambda: fused16(128, 256, 64, False, impl=5), "p256128": lambda: fused16(256, 128, 128, False, impl=5),
    "gnb256": lambda: gnbwd(8, 256, 32), "gnb128": lambda: gnbwd(8, 128, 64), "gnb64": lambda: gnbwd(8, 64, 128), "gnb32": lambda: gnbwd(8, 32, 128),
    "gnb256b32": lambda: gnbwd(32, 256, 32),
    "band32c1": lambda: band(32, False), "band32c2": lambda: band(32, True), "band64c1": lambda: band(64, False),
    "gnstats32": lambda: gnstats(32, 256, DT), "gnstats128": lambda: gnstats(128, 32, torch.float32),
    "gnapply32": lambda: gnapply(32, 256, DT), "gnapply128": lambda: gnapply(128, 32, torch.float32),
    "l1l2": l1l2, "kl": kl, "sample": sample,
    "attn1k": lambda: attn(1024, 128), "attn4k": lambda: attn(4096, 128, 16), "attn4k256": lambda: attn(4096, 256, 8),
    "f32c2sc": lambda: fused_sc(32, 64, 256), "f64c2sc": lambda: fused_sc(64, 32, 128),
    "cout1": cout1, "cout4": cout4, "cin1": cin1, "cin4": cin4,
    "up64": lambda: up(64, 128, True), "up128": lambda: up(128, 64, True), "up128s": lambda: up(128, 32, False),
    "upn64": lambda: up_new(64, 128, True), "upn128": lambda: up_new(128, 64, True), "upn128s": lambda: up_new(128, 32, False),
    "f32c1": lambda impl=0: fused(32, 32, 256, 0, impl=impl), "f32c2": lambda impl=0: fused(32, 32, 256, 1, impl=impl),
    "f64c1": lambda impl=0: fused(64, 64, 128, 0, impl=impl), "f64c2": lambda impl=0: fused(64, 64, 128, 1, impl=impl), "f64c2h": lambda impl=0: fused(64, 64, 128, 1, False, impl=impl),
    "f6432": lambda impl=0: fused(64, 32, 256, 0, impl=impl), "f12864": lambda impl=0: fused(128, 64, 128, 0, impl=impl),
    "f128c1": lambda impl=0: fused(128, 128, 64, 0, impl=impl), "f128c2": lambda impl=0: fused(128, 128, 64, 1, impl=impl),
    "f128c1s": lambda impl=0: fused(128, 128, 32, 0, impl=impl), "f128c2s": lambda impl=0: fused(128, 128, 32, 1, impl=impl),
}
for _k in ("f32c1", "f32c2", "f64c1", "f64c2", "f64c2h", "f6432", "f12864", "f128c1", "f128c2", "f128c1s", "f128c2s"):
    CASES[_k + "@3"] = (lambda f: (lambda: f(impl=3)))(CASES[_k])
for name in (sys.argv[1:] or list(CASES)):
    try:
        ms, by = CASES[name]()
        print(f"{name:10s} {ms:8.4f} ms  {by / ms / 1e6:8.1f} GB/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name:10s} failed: {e}", flush=True)
