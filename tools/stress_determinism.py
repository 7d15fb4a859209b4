"""Stress run: the same batch through reconstruct() many times, every result bit-identical (catches intermittent
races in the hand-rolled mbarrier pipelines).  stress_determinism.py [repeats]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import _pkg
b200 = _pkg.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
# (stream dtype, ...): the fp16 stream routes the 32-wide layers through the row-band kernel (conv_band.cu) when W >= 96
for cfg, b, hw, stream in ((b200.config.AUTOENCODER_DEF_A, 64, 256, torch.float16), (b200.config.AUTOENCODER_DEF_A, 64, 256, torch.float32),
                           (b200.config.AUTOENCODER_DEF_A, 7, 72, torch.float16), (b200.config.AUTOENCODER_DEF_A, 3, 200, torch.float16),
                           (b200.config.AUTOENCODER_DEF_B, 4, 128, torch.float16)):
    torch.manual_seed(1)
    vae = b200.VAEModel.from_config(cfg).cuda().eval()
    vae.autoencoder.set_stream_dtype(stream)
    x = torch.randn(b, 1, hw, hw, device="cuda")
    ref = vae.reconstruct_deterministic(x).clone()
    mu = vae.encode_deterministic(x).clone()
    bad = 0
    for i in range(reps):
        if not torch.equal(vae.reconstruct_deterministic(x), ref) or not torch.equal(vae.encode_deterministic(x), mu):
            bad += 1
    print(f"stream {str(stream).split('.')[-1]} batch {b} {hw}x{hw}: {reps} repeats, {bad} mismatches, finite={bool(torch.isfinite(ref).all())}", flush=True)
    assert bad == 0
print("stress ok")
