"""One row-band conv launch series for ncu / timing: prof_band.py cin res(0/1) [impl]"""
import sys, pathlib, math
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import os
import torch
import _pkg
b200 = _pkg.load(); ops = b200.ops
cin, res = int(sys.argv[1]), int(sys.argv[2])
ops.FUSED_IMPL = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n, hw, cout = 64, 256, 32
SILU = os.environ.get('BAND_NOSILU') is None
x = torch.randn(n, hw, hw, cin, device="cuda", dtype=torch.float16)
wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin), 0, torch.float16)
bias = torch.randn(cout, device="cuda"); ss = torch.randn(n, cin, 2, device="cuda")
r = torch.randn(n, hw, hw, cout, device="cuda").half() if res else None
for _ in range(3):
    ops.conv3x3_fused(x, ss, SILU, wp, bias, residual=r, gn_groups=16, out_f32=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv3x3_fused(x, ss, SILU, wp, bias, residual=r, gn_groups=16, out_f32=False)
e1.record(); torch.cuda.synchronize()
print("ms/launch", e0.elapsed_time(e1) / 10)
if len(sys.argv) > 4:
    lib = b200._lib.lib()
    buf = torch.zeros(4096, device="cuda", dtype=torch.int64)
    lib.ptivae_debug_set_trace(buf.data_ptr())
    ops.conv3x3_fused(x, ss, SILU, wp, bias, residual=r, gn_groups=16, out_f32=False); torch.cuda.synchronize()
    lib.ptivae_debug_set_trace(None)
    t = buf.cpu()
    bt, rt = t[:2048].view(64, 32), t[2048:3072].view(256, 4)
    t0 = int(rt[0, 0])
    print("band: acc_empty | row_ready rl0..5 | issued || epi: start acc_full | u0: ld slot_ready written | u1: ... | stored done")
    for i in range(0, 14):
        print(f"{i:3d} " + " ".join(f"{int(bt[i, k]) - t0:8d}" for k in range(22)))
    print("row: loader_issue  xf_start  xf_done")
    for g in range(0, 60):
        print(f"{g:3d} " + " ".join(f"{(int(rt[g, k]) - t0) if int(rt[g, k]) else -1:8d}" for k in range(4)))
