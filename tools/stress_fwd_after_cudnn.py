"""Do the forward tcgen05/TMA kernels survive cuDNN's TF32 convolution kernels having run in the same process?"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

b200 = _pkg.load()
ops = b200.ops
DEV = "cuda"
F16 = torch.float16
torch.manual_seed(0)
for it in range(6):
    n, h, w, c = 2, 32, 32, 128
    x = torch.randn(n, h, w, c, device=DEV).to(F16)
    wt = (torch.randn(c, c, 3, 3, device=DEV) / 34).to(F16).float()
    bias = torch.zeros(c, device=DEV)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=1).permute(0, 2, 3, 1)   # cuDNN, TF32 allowed
    torch.cuda.synchronize()
    try:
        out = ops.conv_umma(x, ops.pack_conv_weight(wt, 0, F16), bias, 0)
        ss = torch.stack([torch.ones(n, c, device=DEV), torch.zeros(n, c, device=DEV)], dim=-1).contiguous()
        out2 = ops.conv3x3_fused(x, ss, False, ops.pack_conv_weight(wt, 0, F16), bias)
        q = torch.randn(n, 256, 3 * c, device=DEV).to(F16)
        o = ops.attention(q[..., :c], q[..., c:2 * c], q[..., 2 * c:])
        torch.cuda.synchronize()
        e1 = float((out.float() - ref).norm() / ref.norm())
        e2 = float((out2.float() - ref).norm() / ref.norm())
        print(f"it {it}: conv_umma rel {e1:.2e} fused rel {e2:.2e} attn finite {bool(torch.isfinite(o).all())}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"it {it}: EXC {str(ex)[:120]}", flush=True)
        break
