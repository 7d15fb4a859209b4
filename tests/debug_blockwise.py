"""Block-by-block comparison of the CUDA path with the oracle + run-to-run determinism per block."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
from oracle import aekl_ref
b200 = _pkg.load()
from pti_ldm_vae_b200.autoencoderkl import _Act, AEKLResBlock, SpatialAttentionBlock, AEKLDownsample, UpSample
ops = b200.ops
size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fused = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
cfg = b200.config.AUTOENCODER_DEF_A
ref = aekl_ref.seeded_model(cfg, 1234)
vae = b200.VAEModel.from_config(cfg); vae.load_state_dict(ref.state_dict()); vae = vae.cuda().eval()
ae = vae.autoencoder; ex = ae._exec; ex.fused_stats = fused
x = aekl_ref.synthetic_images(2, size, size, seed=0)

def rel(a, b): return float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
def nchw(t): return t.float().permute(0, 3, 1, 2)

def run_stack(blocks, rblocks, xin, tag):
    with torch.no_grad():
        hr = rblocks[0](xin)
    def first(): return _Act(ops.conv3x3_small_cin(xin.cuda(), blocks[0].conv.weight.detach(), blocks[0].conv.bias.detach(), ))
    a = first()
    print(f"{tag}.0 conv_in rel {rel(nchw(a.t), hr):.3e}")
    for i, (blk, rblk) in enumerate(zip(list(blocks)[1:-2], list(rblocks)[1:-2]), start=1):
        with torch.no_grad():
            hr_next = rblk(hr)
        # feed OUR block the oracle's input (bf16-rounded) so errors do not accumulate
        ain = _Act(hr.permute(0, 2, 3, 1).contiguous().cuda())
        def runblk(inp):
            inp = _Act(inp.t)
            if isinstance(blk, AEKLResBlock): return ex.resblock(blk, inp, True, True)
            if isinstance(blk, SpatialAttentionBlock): return ex.attention(blk, inp, True, True)
            if isinstance(blk, AEKLDownsample): return ex.conv(inp.t.half(), blk.conv.conv, 1, out_f32=True)
            return ex.conv(inp.t.half(), blk.postconv.conv, 2, out_f32=True)
        o1 = runblk(ain); o2 = runblk(ain)
        torch.cuda.synchronize()
        same = torch.equal(o1.t, o2.t)
        chained = runblk(a)
        print(f"{tag}.{i} {type(blk).__name__:22s} isolated rel {rel(nchw(o1.t), hr_next):.3e}  chained rel {rel(nchw(chained.t), hr_next):.3e}  deterministic={same}  maxdiff2runs={float((o1.t.float()-o2.t.float()).abs().max()):.3e}")
        a = chained; hr = hr_next
    with torch.no_grad():
        hr_f = rblocks[-1](rblocks[-2](hr))
    ss = ex.scale_shift(a, blocks[-2])
    out = ops.conv3x3_small_cout(a.t, blocks[-1].conv.weight.detach(), blocks[-1].conv.bias.detach(), ss)
    print(f"{tag}.final rel {rel(out, hr_f):.3e}")
    return out, hr_f

h, hr = run_stack(ae.encoder.blocks, ref.encoder.blocks, x, "enc")
with torch.no_grad():
    mu_r, sg_r = ref.encode(x)
mu, sg = ae.encode(x.cuda())
print("z_mu rel", rel(mu, mu_r), "z_sigma rel", rel(sg, sg_r))
with torch.no_grad():
    zq_r = ref.post_quant_conv(mu_r)
run_stack(ae.decoder.blocks, ref.decoder.blocks, zq_r, "dec")
