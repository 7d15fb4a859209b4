"""Fused-conv micro-benchmark + CTA-0 role timeline.  usage: prof_fused.py [trace]"""
import sys, pathlib, math
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
b200 = _pkg.load(); ops = b200.ops; lib = b200._lib.lib()
trace = len(sys.argv) > 1 and sys.argv[1] == "trace"
if len(sys.argv) > 2: ops.FUSED_IMPL = int(sys.argv[2])
print("FUSED_IMPL", ops.FUSED_IMPL)
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
H16 = len(sys.argv) > 4 and sys.argv[4] == "h16"    # 16-bit residual stream: 16-bit in / residual / out
cases = [(64, 256, 256, 32, 32, True), (64, 256, 256, 32, 32, False), (64, 64, 64, 128, 128, True), (64, 128, 128, 64, 64, True), (64, 128, 128, 64, 64, False), (64, 256, 256, 64, 32, False), (64, 128, 128, 32, 64, False)]
cases += [(64, 64, 64, 128, 128, False), (64, 128, 128, 128, 64, False), (64, 32, 32, 128, 128, True)]
for (n, h, w, cin, cout, conv2) in cases:
    if only and f"{cin}-{cout}-{h}-{int(conv2)}" not in only:
        continue
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, cin, device="cuda", dtype=torch.float32 if not (conv2 or H16) else torch.float16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).cuda()
    wp = ops.pack_conv_weight(wt, 0, torch.float16)
    bias = torch.randn(cout, generator=g).cuda()
    ss = torch.randn(n, cin, 2, device="cuda")
    res = torch.randn(n, h, w, cout, device="cuda") if conv2 else None
    if H16 and conv2:
        res = res.half()
    def run():
        return ops.conv3x3_fused(x, ss, True, wp, bias, residual=res, gn_groups=16, out_f32=conv2 and not H16)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * h * w * cin * cout * 9
    by = n * h * w * (x.element_size() * cin + (4 if conv2 else 2) * cout + (4 * cout if conv2 else 0))
    print(f"{cin}->{cout}@{h} conv{'2' if conv2 else '1'}: {ms:.3f} ms  {fl/ms/1e9:.0f} TF  {by/ms/1e6:.0f} GB/s", flush=True)
    if trace:
        buf = torch.zeros(64 * 32, device="cuda", dtype=torch.int64)
        lib.ptivae_debug_set_trace(buf.data_ptr())
        run(); torch.cuda.synchronize()
        lib.ptivae_debug_set_trace(None)
        t = buf.view(64, 32).cpu()
        t0 = int(t[0, 0])
        names = ["xf_start", "xf_end", "mma_start", "mma_issued", "epi_start", "epi_end"]
        print("   tile " + " ".join(f"{nm:>10s}" for nm in names))
        for i in range(0, 7):
            print(f"   {i:4d} " + " ".join(f"{int(t[i, k]) - t0:10d}" for k in range(6)) + "   | os_ready acc_ready drained: " + " ".join(f"{int(t[i, k]) - int(t[i, 4]):7d}" for k in (6, 7, 8, 9, 10)))
        for i in (3, 4):
            x0 = int(t[i, 0])
            print(f"   tile {i} transform pieces (rel. to xf_start) start: " + " ".join(str(int(t[i, 14 + k]) - x0) for k in range(4)) + "  end: " + " ".join(str(int(t[i, 18 + k]) - x0) for k in range(4)) + "  | mma chunk ready (rel. mma_start): " + " ".join(str(int(t[i, 22 + k]) - int(t[i, 2])) for k in range(2)))
        for i in (3, 4) if ops.FUSED_IMPL != 3 else ():
            m0 = int(t[i, 2])
            print(f"   tile {i} mma detail (rel. to mma_start): wait-done " + " ".join(str(int(t[i, 14 + k]) - m0) for k in range(9)))
            print(f"   tile {i}                               committed " + " ".join(str(int(t[i, 23 + k]) - m0) for k in range(9)))
