"""One fused conv shape for ncu: prof_one.py cin cout hw conv2(0/1)"""
import sys, pathlib, math
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
b200 = _pkg.load(); ops = b200.ops
cin, cout, hw, conv2 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = 64
x = torch.randn(n, hw, hw, cin, device="cuda", dtype=torch.float16 if conv2 else torch.float32)
wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(9 * cin), 0, torch.float16)
bias = torch.randn(cout, device="cuda"); ss = torch.randn(n, cin, 2, device="cuda")
res = torch.randn(n, hw, hw, cout, device="cuda") if conv2 else None
for _ in range(3):
    ops.conv3x3_fused(x, ss, True, wp, bias, residual=res, gn_groups=16, out_f32=bool(conv2))
torch.cuda.synchronize()
print("done")
