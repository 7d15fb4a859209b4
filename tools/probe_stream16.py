"""fp32 vs fp16 residual stream: parity against the CPU oracle and CUDA-graph step time (config A, B=64, 256^2)."""
import json
import sys
import pathlib
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import _pkg  # noqa: E402
import bench  # noqa: E402
from oracle import aekl_ref  # noqa: E402

b200 = _pkg.load()
dev = torch.device("cuda:0")
out = {}
for cfgname in ("AUTOENCODER_DEF_A", "AUTOENCODER_DEF_B"):
    cfg = getattr(b200.config, cfgname)
    ref = aekl_ref.seeded_model(cfg, 1234)
    vae = b200.VAEModel.from_config(cfg)
    vae.load_state_dict(ref.state_dict(), strict=True)
    vae = vae.to(dev).eval()
    x = aekl_ref.synthetic_images(2, 128, 128, seed=0)
    with torch.no_grad():
        mu_r, sg_r = ref.encode(x)
        eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(7))
        rec_r, _, _ = ref(x, eps)
    rl = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())  # noqa: E731
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
        vae.autoencoder.set_stream_dtype(dt)
        rec, mu, sg = vae.autoencoder(x.to(dev), eps.to(dev))
        out[f"{cfgname}_{name}_parity"] = {"recon": rl(rec, rec_r), "z_mu": rl(mu, mu_r), "z_sigma": rl(sg, sg_r)}
    if cfgname.endswith("A"):
        B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
        xd = aekl_ref.synthetic_images(B, 256, 256, seed=0).to(dev)
        for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
            vae.autoencoder.set_stream_dtype(dt)
            g = b200.GraphedVAE(vae, B, 256, 256, mode="forward", warmup=2)
            g.x.copy_(xd)
            for _ in range(3):
                g()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                g()
            e1.record()
            torch.cuda.synchronize()
            out[f"A_{name}_ms_per_step"] = e0.elapsed_time(e1) / 20
            vae.autoencoder._rng_dev = None
            agg = bench.kernel_breakdown(vae.autoencoder, xd, passes=2)
            out[f"A_{name}_breakdown"] = {k: [round(v["ms"] / 2, 4), v["launches"] // 2, round(v["bytes"] / (v["ms"] * 1e-3) / 1e9),
                                              round(v["flops"] / (v["ms"] * 1e-3) / 1e12)]
                                          for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
            out[f"A_{name}_eager_ms"] = sum(v["ms"] for v in agg.values()) / 2
print(json.dumps(out, indent=1))
