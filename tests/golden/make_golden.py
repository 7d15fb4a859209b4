"""Generates the committed golden fixtures (run in the build container, NOT on the GPU box):

  python tests/golden/make_golden.py

* ``losses_ref.npz``  -- outputs of the REFERENCE's own ``compute_kl_loss`` / ``compute_total_loss``
  (/root/reference/src/pti_ldm_vae/models/losses.py, imported by file path: it only needs torch) on
  seeded inputs.  This is the one piece of the hot path the reference itself can execute here; it
  pins oracle/aekl_ref.py::kl_loss_ref / total_loss_ref.
* ``aekl_*.npz``      -- outputs of the oracle restatement of monai 1.5.1 AutoencoderKL (MONAI is
  not installable here -> "parity unpinned" for the network body) on seeded weights/inputs/eps, in
  fp32 and fp64, with a few intermediate activations for bisecting.
"""
import importlib.util
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import aekl_ref  # noqa: E402

import _pkg  # noqa: E402

CFG = _pkg.load().config


def ref_losses():
    p = pathlib.Path("/root/reference/src/pti_ldm_vae/models/losses.py")
    spec = importlib.util.spec_from_file_location("ref_losses", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(42)
    mu = torch.randn(6, 4, 32, 32, generator=g)
    sigma = torch.rand(6, 4, 32, 32, generator=g) * 1.5 + 0.05
    out = {
        "mu": mu.numpy(), "sigma": sigma.numpy(),
        "kl_as_called": mod.compute_kl_loss(mu, sigma).numpy(),
        "kl_sigma_mode": mod.compute_kl_loss(mu, sigma, input_is_logvar=False).numpy(),
    }
    t = [torch.tensor(v) for v in (0.31, 12.5, 0.77, 0.2, 1.9)]
    out["total_ar"] = mod.compute_total_loss(*t, kl_weight=1e-4, perceptual_weight=1.0, adv_weight=0.5,
                                             ar_gamma=0.5, ar_vae_enabled=True).numpy()
    out["total_noar"] = mod.compute_total_loss(*t, kl_weight=1e-3, perceptual_weight=1.0, adv_weight=3.0,
                                               ar_gamma=0.5, ar_vae_enabled=False).numpy()
    # AR-VAE loss of the reference itself ("all" pairs and a seeded "subset")
    import random
    zb = torch.randn(12, 10, generator=g)
    attrs = {f"a{k}": torch.randint(20, 200, (12,), generator=g).float() for k in range(3)}
    attrs["a1"][3] = attrs["a1"][7]              # ties -> sign 0 pairs are dropped
    mapping = {"a0": {"latent_channel": 0, "delta": 1.0}, "a1": {"latent_channel": 4}, "a2": {"latent_channel": 9, "delta": 0.25}}
    dg = {"enabled": True, "value": 2.0}
    tot, per, cnt, dl = mod.compute_ar_vae_loss(zb, attrs, mapping, "all", None, dg)
    out.update({"ar_z": zb.numpy(), "ar_total": tot.numpy(), "ar_per": np.array([float(per[k]) for k in mapping]),
                "ar_cnt": np.array([cnt[k] for k in mapping]), "ar_delta": np.array([dl[k] for k in mapping])})
    for k, v in attrs.items():
        out["ar_attr_" + k] = v.numpy()
    random.seed(123)
    tot, per, cnt, _ = mod.compute_ar_vae_loss(zb, attrs, mapping, "subset", 40, dg)
    out.update({"ars_total": tot.numpy(), "ars_per": np.array([float(per[k]) for k in mapping]),
                "ars_cnt": np.array([cnt[k] for k in mapping])})
    z4 = torch.randn(5, 10, 4, 4, generator=g)
    tot4, _, _, _ = mod.compute_ar_vae_loss(z4, {k: v[:5] for k, v in attrs.items()}, mapping, "all", None, dg)
    out.update({"ar_z4": z4.numpy(), "ar_total4": tot4.numpy()})
    np.savez_compressed(HERE / "losses_ref.npz", **out)
    print("losses_ref.npz", {k: float(v) for k, v in out.items() if v.ndim == 0})


def ref_regressor():
    """LatentRegressor of the reference itself (regression_head.py imports pti_ldm_vae.models.autoencoder, which
    needs MONAI: stub that module, the class under test is pure torch)."""
    import types
    stub = types.ModuleType("pti_ldm_vae.models.autoencoder")
    stub.VAEModel = object
    for name in ("pti_ldm_vae", "pti_ldm_vae.models"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pti_ldm_vae.models.autoencoder"] = stub
    p = pathlib.Path("/root/reference/src/pti_ldm_vae/models/regression_head.py")
    spec = importlib.util.spec_from_file_location("ref_reg", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for tag, act, drop in (("relu", "relu", 0.1), ("gelu", "gelu", 0.0), ("lrelu", "leaky_relu", 0.0), ("elu", "elu", 0.0)):
        torch.manual_seed(77)
        reg = mod.LatentRegressor(in_features=4096, hidden_dims=[256, 32], output_dim=6, dropout=drop, activation=act).eval()
        x = torch.randn(5, 4096, generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            out[tag] = reg(x).numpy()
        out[tag + "_keys"] = np.array(list(reg.state_dict().keys()))
    out["x"] = x.numpy()
    np.savez_compressed(HERE / "regressor_ref.npz", **out)
    print("regressor_ref.npz", {k: v.shape for k, v in out.items() if not k.endswith("_keys")})


def param_checksum(model):
    return float(sum(p.detach().double().abs().sum() for p in model.parameters()))


def aekl_case(name, cfg, b, h, w, taps):
    model = aekl_ref.seeded_model(cfg, 1234)
    x = aekl_ref.synthetic_images(b, h, w, seed=0)
    with torch.no_grad():
        mu, sigma = model.encode(x)
        eps = torch.randn(mu.shape, generator=torch.Generator().manual_seed(7))
        recon, mu2, sigma2 = model(x, eps)
        assert torch.equal(mu, mu2)
        inter = {}
        hcur = x
        for i, blk in enumerate(model.encoder.blocks):
            hcur = blk(hcur)
            if f"encoder.blocks.{i}" in taps:
                inter[f"encoder.blocks.{i}"] = hcur.numpy()
        hcur = model.post_quant_conv(model.sampling(mu, sigma, eps))
        for i, blk in enumerate(model.decoder.blocks):
            hcur = blk(hcur)
            if f"decoder.blocks.{i}" in taps:
                inter[f"decoder.blocks.{i}"] = hcur.numpy()
        m64 = aekl_ref.seeded_model(cfg, 1234).double()
        recon64, mu64, sigma64 = m64(x.double(), eps.double())
        rdet = model.reconstruct(x)
    out = dict(eps=eps.numpy(), recon=recon.numpy(), z_mu=mu.numpy(), z_sigma=sigma.numpy(),
               recon_det=rdet.numpy(), recon64=recon64.float().numpy(), z_mu64=mu64.float().numpy(),
               z_sigma64=sigma64.float().numpy(),
               kl_as_called=aekl_ref.kl_loss_ref(mu, sigma).numpy(),
               kl_sigma_mode=aekl_ref.kl_loss_ref(mu, sigma, input_is_logvar=False).numpy(),
               l1=aekl_ref.l1_ref(recon, x).numpy(), l2=aekl_ref.l2_ref(recon, x).numpy(),
               param_checksum=np.float64(param_checksum(model)), x_checksum=np.float64(x.double().abs().sum()))
    out.update({"tap/" + k: v for k, v in inter.items()})
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, "recon", tuple(recon.shape), "fp32-vs-fp64 rel-L2",
          float((recon - recon64.float()).norm() / recon64.float().norm()))


def ref_metrics():
    """Outputs of the reference's OWN eval metrics (utils/eval_metrics.py) and LocalNormalizeByMask
    (data/transforms.py; tifffile stubbed, it is only used by the TIFF reader) -> metrics_ref.npz."""
    import types
    spec = importlib.util.spec_from_file_location("ref_evalm", "/root/reference/src/pti_ldm_vae/utils/eval_metrics.py")
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    sys.modules.setdefault("tifffile", types.ModuleType("tifffile"))
    spec = importlib.util.spec_from_file_location("ref_tf", "/root/reference/src/pti_ldm_vae/data/transforms.py")
    tf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tf)
    g = torch.Generator().manual_seed(21)
    out = {}
    for tag, (b, c, h, w) in {"a": (3, 1, 70, 50), "b": (2, 1, 256, 256), "c": (4, 1, 33, 64)}.items():
        img = torch.rand(b, c, h, w, generator=g) * 1.2 - 0.1            # some values outside [0, 1]
        img[..., : w // 4] = 0.0
        rec = img + 0.05 * torch.randn(b, c, h, w, generator=g)
        rc, ic = torch.clamp(rec, 0.0, 1.0), torch.clamp(img, 0.0, 1.0)
        out[f"{tag}_img"], out[f"{tag}_rec"] = img.numpy(), rec.numpy()
        out[f"{tag}_psnr"] = em.compute_psnr(rc, ic).numpy()
        out[f"{tag}_ssim"] = em.compute_ssim(rc, ic).numpy()
        out[f"{tag}_mse"] = torch.mean((rc - ic) ** 2, dim=(1, 2, 3)).numpy()
        out[f"{tag}_mae"] = torch.mean(torch.abs(rc - ic), dim=(1, 2, 3)).numpy()
        out[f"{tag}_psnr_raw"] = em.compute_psnr(rec, img, data_range=2.0).numpy()
        out[f"{tag}_ssim_raw"] = em.compute_ssim(rec, img, data_range=2.0, k1=0.02, k2=0.05).numpy()
    norm = tf.LocalNormalizeByMask()
    raw = torch.rand(5, 1, 96, 80, generator=g) * 200.0 + 20.0
    raw[:, :, :, :30] = 0.0                                               # background band
    raw[1, :, 40:, :] = 0.0
    raw[3] = 0.0
    raw[3, 0, 10:20, 40:50] = 7.0                                         # constant foreground: std <= 1e-5 -> 1
    out["ln_raw"] = raw.numpy()
    out["ln_out"] = np.stack([norm(raw[i].numpy().copy()) for i in range(5) if i != 4] + [norm(raw[4].clone())])
    np.savez_compressed(HERE / "metrics_ref.npz", **out)
    print("metrics_ref: ssim", out["a_ssim"], "psnr", out["a_psnr"])


if __name__ == "__main__":
    if pathlib.Path("/root/reference").exists():
        ref_losses()
        ref_regressor()
        ref_metrics()
    aekl_case("aekl_A_64", CFG.AUTOENCODER_DEF_A, 2, 64, 64, {"encoder.blocks.3", "encoder.blocks.13", "decoder.blocks.6"})
    aekl_case("aekl_A_256", CFG.AUTOENCODER_DEF_A, 1, 256, 256, set())
    aekl_case("aekl_B_64", CFG.AUTOENCODER_DEF_B, 1, 64, 64, {"encoder.blocks.10"})
