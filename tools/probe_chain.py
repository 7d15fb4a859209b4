"""Chained-launch A/B: config A forward (batch 64, 256x256, fp16 stream), graph replay and eager, under the current
PTIVAE_CHAIN / PTIVAE_LIB environment.  probe_chain.py [label]"""
import sys, pathlib, os
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
b200 = _pkg.load()
torch.manual_seed(1234)
vae = b200.VAEModel.from_config(b200.config.AUTOENCODER_DEF_A).cuda().eval()
vae.autoencoder.set_stream_dtype(torch.float16)
g = b200.GraphedVAE(vae, 64, 256, 256, mode="forward")
g.x.copy_(torch.randn(64, 1, 256, 256, generator=torch.Generator().manual_seed(0)).cuda())


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x = g.x.clone()
with torch.no_grad():
    print(f"{sys.argv[1] if len(sys.argv) > 1 else '':24s} chain={os.environ.get('PTIVAE_CHAIN', 'default')} graph {timeit(g):.3f} ms  eager {timeit(lambda: vae(x)):.3f} ms",
          flush=True)
