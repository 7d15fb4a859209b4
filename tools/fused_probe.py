"""Which shifted-descriptor variant does the hardware want?  Runs the fused conv test cases with
base_offset = 0 and = (start>>7)&7 and prints the error of each."""
import sys, pathlib, math
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
import _pkg
b200 = _pkg.load(); ops = b200.ops
torch.backends.cudnn.allow_tf32 = False
for variant in (0, 1):
    ops.FUSED_DESC_BASE_OFFSET = variant
    for (n, h, w, cin, cout) in [(1, 16, 16, 64, 64), (2, 32, 32, 128, 128), (1, 32, 32, 32, 32), (2, 40, 24, 64, 128)]:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, h, w, cin, generator=g).cuda()
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).cuda().half().float()
        bias = torch.randn(cout, generator=g).cuda()
        ref = F.conv2d(x.half().float().permute(0, 3, 1, 2), wt, bias, padding=1).permute(0, 2, 3, 1)
        try:
            out = ops.conv3x3_fused(x, None, False, ops.pack_conv_weight(wt, 0, torch.float16), bias, out_f32=True)
            torch.cuda.synchronize()
            rel = float((out - ref).norm() / ref.norm())
        except Exception as e:  # noqa
            rel = f"EXC {e}"
        print(f"variant base_offset={variant} {cin}->{cout} {h}x{w} n={n}: rel {rel}", flush=True)
