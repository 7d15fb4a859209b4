"""End-to-end parity of the CUDA path (through VAEModel -> C ABI) against the CPU oracle and the
committed golden fixtures, on identical weights / inputs / eps.

Tolerances (BASELINE.json north_star): rel-L2 <= 1e-2 on recon and <= 5e-3 on z_mu/z_sigma; KL and
reconstruction losses within 1e-3 relative.  The default operand format (fp16 tensor-core operands, fp32
accumulate, fp32 residual stream) must meet them.  The bf16 operand format is also exercised: its
per-block operand rounding (~2.5e-3, measured) accumulates over the 14+14 blocks to ~1e-2 at z_mu and
~2e-2 at recon (measured 0.9e-2 / 1.8-2.1e-2), so it is a secondary mode held to 4x the tolerances
(documented in DESIGN.md section 3.5) -- it cannot meet the north-star numbers, fp16 can.
"""
import pathlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).resolve().parent / "golden"
DEV = "cuda"
TOL_RECON, TOL_LATENT, TOL_LOSS = 1e-2, 5e-3, 1e-3


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def _models(b200, oracle, cfg):
    ref = oracle.seeded_model(cfg, 1234)
    vae = b200.VAEModel.from_config(cfg)
    vae.load_state_dict(ref.state_dict(), strict=True)
    return ref, vae.to(DEV).eval()


@pytest.mark.parametrize("name,cfgname,b,h,w", [("aekl_A_64", "AUTOENCODER_DEF_A", 2, 64, 64),
                                               ("aekl_A_256", "AUTOENCODER_DEF_A", 1, 256, 256),
                                               ("aekl_B_64", "AUTOENCODER_DEF_B", 1, 64, 64)])
@pytest.mark.parametrize("fused_stats,op_dtype,fused_conv", [(True, torch.float16, True), (False, torch.float16, True),
                                                            (True, torch.bfloat16, True), (True, torch.float16, False)])
def test_forward_matches_golden(b200, oracle, name, cfgname, b, h, w, fused_stats, op_dtype, fused_conv):
    cfg = getattr(b200.config, cfgname)
    gold = np.load(GOLD / f"{name}.npz")
    ref, vae = _models(b200, oracle, cfg)
    assert abs(sum(float(p.detach().double().abs().sum()) for p in ref.parameters()) - float(gold["param_checksum"])) < 1e-6 * float(gold["param_checksum"])
    x = oracle.synthetic_images(b, h, w, seed=0)
    assert abs(float(x.double().abs().sum()) - float(gold["x_checksum"])) < 1e-9 * float(gold["x_checksum"])
    eps = torch.from_numpy(gold["eps"])
    vae.autoencoder.set_fused_stats(fused_stats)
    vae.autoencoder.set_operand_dtype(op_dtype)
    vae.autoencoder.set_fused_conv(fused_conv)
    tol_latent = TOL_LATENT if op_dtype == torch.float16 else 4 * TOL_LATENT
    tol_recon = TOL_RECON if op_dtype == torch.float16 else 4 * TOL_RECON
    tol_loss = TOL_LOSS if op_dtype == torch.float16 else 5 * TOL_LOSS
    recon, mu, sigma = vae.autoencoder(x.to(DEV), eps.to(DEV))
    assert recon.shape == x.shape and recon.dtype == torch.float32
    e_mu = _rel_l2(mu, torch.from_numpy(gold["z_mu64"]))
    e_sg = _rel_l2(sigma, torch.from_numpy(gold["z_sigma64"]))
    e_rc = _rel_l2(recon, torch.from_numpy(gold["recon64"]))
    print(f"{name} fused_stats={fused_stats} fused_conv={fused_conv} {op_dtype}: rel-L2 recon {e_rc:.2e} z_mu {e_mu:.2e} z_sigma {e_sg:.2e}")
    assert e_mu <= tol_latent and e_sg <= tol_latent, (e_mu, e_sg)
    assert e_rc <= tol_recon, e_rc
    # losses through the product's own reductions
    kl = float(b200.compute_kl_loss(mu, sigma))
    l1 = float(b200.l1_loss(recon, x.to(DEV)))
    l2 = float(b200.mse_loss(recon, x.to(DEV)))
    print(f"   kl {kl:.6g} vs {float(gold['kl_as_called']):.6g}  l1 {l1:.6g} vs {float(gold['l1']):.6g}  l2 {l2:.6g} vs {float(gold['l2']):.6g}")
    assert abs(kl - float(gold["kl_as_called"])) <= tol_loss * abs(float(gold["kl_as_called"]))
    assert abs(l1 - float(gold["l1"])) <= tol_loss * float(gold["l1"])
    assert abs(l2 - float(gold["l2"])) <= 2 * tol_loss * float(gold["l2"])
    # deterministic path (inference_vae.py:78)
    rdet = vae.reconstruct_deterministic(x.to(DEV))
    assert _rel_l2(rdet, torch.from_numpy(gold["recon_det"])) <= tol_recon


@pytest.mark.parametrize("b,h,w", [(3, 96, 160), (2, 40, 56), (5, 8, 8), (1, 136, 24)])
def test_forward_matches_oracle_live(b200, oracle, b, h, w):
    """Fresh seeds (not the golden ones), extents whose pyramid levels are not tile multiples (40x56 -> 5x7 at the
    bottom; 8x8 -> a 1x1 latent), oracle evaluated in this process."""
    cfg = b200.config.AUTOENCODER_DEF_A
    ref, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(b, h, w, seed=11)
    with torch.no_grad():
        mu_r, sg_r = ref.encode(x)
        eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(3))
        rec_r, _, _ = ref(x, eps)
    rec, mu, sg = vae.autoencoder(x.to(DEV), eps.to(DEV))
    assert _rel_l2(mu, mu_r) <= TOL_LATENT and _rel_l2(sg, sg_r) <= TOL_LATENT
    assert _rel_l2(rec, rec_r) <= TOL_RECON
    z = vae.encode_stage_2_inputs(x.to(DEV))
    assert z.shape == mu_r.shape and torch.isfinite(z).all()
    dec = vae.decode_stage_2_outputs(mu_r.to(DEV))
    with torch.no_grad():
        assert _rel_l2(dec, ref.decode(mu_r)) <= TOL_RECON


@pytest.mark.parametrize("cfg_name", ["A", "B"])
def test_chained_launch_masks_are_bit_identical(b200, cfg_name):
    """Programmatic dependent launches (ptivae_set_chained_launch) only move WHEN a kernel's set-up runs: the forward of
    both configurations, eager and as a replayed CUDA graph, is bit-identical for every mask, on both residual streams."""
    cfg = b200.config.AUTOENCODER_DEF_A if cfg_name == "A" else b200.config.AUTOENCODER_DEF_B
    torch.manual_seed(5)
    vae = b200.VAEModel.from_config(cfg).to(DEV).eval()
    x = torch.randn(3, 1, 128, 144, device=DEV)
    was = b200.ops.set_chained_launch(0)
    try:
        for stream in (torch.float16, torch.float32):
            vae.autoencoder.set_stream_dtype(stream)
            b200.ops.set_chained_launch(0)
            ref = (vae.reconstruct_deterministic(x).clone(), vae.encode_deterministic(x).clone())
            for mask in (1, 2, 3):
                b200.ops.set_chained_launch(mask)
                for _ in range(3):
                    assert torch.equal(vae.reconstruct_deterministic(x), ref[0]), (stream, mask)
                    assert torch.equal(vae.encode_deterministic(x), ref[1]), (stream, mask)
                g = b200.GraphedVAE(vae, 3, 128, 144, mode="reconstruct")
                g.x.copy_(x)
                for _ in range(3):
                    out = g()
                    out = out[0] if isinstance(out, (tuple, list)) else out
                    assert torch.equal(out, ref[0]), (stream, mask, "graph")
    finally:
        b200.ops.set_chained_launch(was)


@pytest.mark.parametrize("h,w", [(64, 64), (128, 144)])
def test_fp16_stream_matches_oracle(b200, oracle, h, w):
    """Inference with the residual stream in the fp16 operand format (AutoencoderKL.set_stream_dtype; the row-band conv
    kernel takes the 32-output-channel layers when W >= 96): same gates against the fp32 CPU oracle as the fp32 stream."""
    for cfgname in ("AUTOENCODER_DEF_A", "AUTOENCODER_DEF_B"):
        ref, vae = _models(b200, oracle, getattr(b200.config, cfgname))
        x = oracle.synthetic_images(2, h, w, seed=4)
        with torch.no_grad():
            mu_r, sg_r = ref.encode(x)
            eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(7))
            rec_r, _, _ = ref(x, eps)
        vae.autoencoder.set_stream_dtype(torch.float16)
        rec, mu, sg = vae.autoencoder(x.to(DEV), eps.to(DEV))
        assert _rel_l2(rec, rec_r) <= 1e-2 and _rel_l2(mu, mu_r) <= 5e-3 and _rel_l2(sg, sg_r) <= 5e-3, cfgname
        rec2, _, _ = vae.autoencoder(x.to(DEV), eps.to(DEV))
        assert torch.equal(rec, rec2), "deterministic"
        assert vae.autoencoder.check_range(x.to(DEV))
        with pytest.raises(ValueError):
            vae.autoencoder.set_operand_dtype(torch.bfloat16)       # a bf16 stream would miss the tolerances
        vae.autoencoder.set_stream_dtype(torch.float32)
        vae.autoencoder.set_operand_dtype(torch.bfloat16)
        with pytest.raises(ValueError):
            vae.autoencoder.set_stream_dtype(torch.float16)


@pytest.mark.parametrize("stream", [torch.float32, torch.float16])
def test_fp16_range_large_streams(b200, oracle, stream):
    """fp16 stores clamp at +-65504 (the reference is fp32).  A ResBlock whose conv2 is scaled so that
    the residual stream sits at ~3e3 (rms; the check is conservative by sqrt(tile elements)) must still meet the gates (the clamp is far away, relative precision is unchanged)
    and check_range must say so; scaled to ~3e5 check_range must flag it, and the documented route for such a checkpoint
    (fp32 stream + bf16 operands) must stay finite and within the bf16 tolerances."""
    cfg = b200.config.AUTOENCODER_DEF_A
    x = oracle.synthetic_images(1, 64, 128, seed=6)
    for scale, expect_ok in ((1.0e3, True), (1.0e5, False)):
        ref = oracle.seeded_model(cfg, 1234)
        with torch.no_grad():
            ref.encoder.blocks[1].conv2.conv.weight.mul_(scale)
            ref.encoder.blocks[1].conv2.conv.bias.mul_(scale)
            ref.decoder.blocks[13].conv2.conv.weight.mul_(scale)
        vae = b200.VAEModel.from_config(cfg)
        vae.load_state_dict(ref.state_dict(), strict=True)
        vae = vae.to(DEV).eval()
        with torch.no_grad():
            mu_r, _ = ref.encode(x)
            rec_r = ref.decode(mu_r)
        ae = vae.autoencoder
        ae.set_stream_dtype(stream)
        assert ae.check_range(x.to(DEV)) == expect_ok, (scale, stream)
        if expect_ok:
            mu = ae.encode(x.to(DEV))[0]
            rec = ae.decode(mu)
            assert torch.isfinite(rec).all()
            assert _rel_l2(mu, mu_r) <= 5e-3 and _rel_l2(rec, rec_r) <= 1e-2, (scale, stream)
        else:
            ae.set_stream_dtype(torch.float32)
            ae.set_operand_dtype(torch.bfloat16)
            mu = ae.encode(x.to(DEV))[0]
            rec = ae.decode(mu)
            assert torch.isfinite(rec).all() and torch.isfinite(mu).all()
            assert _rel_l2(mu, mu_r) <= 4 * 5e-3 and _rel_l2(rec, rec_r) <= 4 * 1e-2, (scale, stream)


def test_api_contract(b200, oracle):
    cfg = b200.config.AUTOENCODER_DEF_A
    vae = b200.VAEModel.from_config(cfg).to(DEV).eval()
    x = oracle.synthetic_images(2, 64, 64).to(DEV)
    out = vae(x)
    assert isinstance(out, tuple) and len(out) == 3
    recon, mu, sigma = out
    assert mu.shape == (2, 4, 8, 8) and sigma.shape == mu.shape and (sigma > 0).all()
    # sampling is stochastic in eval mode too (reference semantics), deterministic encode is not
    r2, mu2, _ = vae(x)
    assert torch.equal(mu, mu2) and not torch.equal(recon, r2)
    assert torch.equal(vae.encode_deterministic(x), mu)
    # linearity-free sanity at full size: batch independence (GroupNorm/attention are per-sample)
    xb = oracle.synthetic_images(4, 64, 64, seed=5).to(DEV)
    m_all = vae.encode_deterministic(xb)
    m_one = vae.encode_deterministic(xb[2:3])
    assert torch.allclose(m_all[2:3], m_one, rtol=1e-3, atol=1e-4)
    with pytest.raises(RuntimeError):
        vae(x.cpu())
    # train mode: forward is differentiable (train_vae.py:385,444), the stand-alone encode()/decode() are inference-only
    vae.train()
    rec_t, mu_t, _ = vae(x)
    assert rec_t.requires_grad and mu_t.requires_grad
    rec_t.float().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in vae.parameters())
    with pytest.raises(NotImplementedError):
        vae.autoencoder.encode(x)


def test_graph_replay_matches_eager(b200, oracle):
    cfg = b200.config.AUTOENCODER_DEF_A
    ref, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(2, 64, 64, seed=2).to(DEV)
    eager = vae.reconstruct_deterministic(x)
    g = b200.GraphedVAE(vae, 2, 64, 64, mode="reconstruct")
    out = g(x)
    torch.cuda.synchronize()
    assert _rel_l2(out, eager) <= 1e-3
    gf = b200.GraphedVAE(vae, 2, 64, 64, mode="forward")
    r1 = gf(x)[0].clone()
    r2 = gf(x)[0].clone()
    torch.cuda.synchronize()
    assert not torch.equal(r1, r2), "noise must be fresh on every replay"
    vae.autoencoder._rng_dev = None


AR_MAPPING = {"a0": {"latent_channel": 0, "delta": 1.0}, "a1": {"latent_channel": 4}, "a2": {"latent_channel": 9, "delta": 0.25}}
AR_DG = {"enabled": True, "value": 2.0}


def test_ar_vae_loss_matches_reference_golden(b200, oracle):
    """compute_ar_vae_loss (row a17) vs outputs of the reference's own function (tests/golden/losses_ref.npz)."""
    import random
    g = np.load(GOLD / "losses_ref.npz")
    z = torch.from_numpy(g["ar_z"]).to(DEV)
    attrs = {k: torch.from_numpy(g["ar_attr_" + k]) for k in AR_MAPPING}          # CPU tensors, as a DataLoader yields
    tot, per, cnt, dl = b200.compute_ar_vae_loss(z, attrs, AR_MAPPING, "all", None, AR_DG)
    assert abs(float(tot) - float(g["ar_total"])) <= TOL_LOSS * float(g["ar_total"])
    assert np.allclose([float(per[k]) for k in AR_MAPPING], g["ar_per"], rtol=TOL_LOSS)
    assert [cnt[k] for k in AR_MAPPING] == list(g["ar_cnt"]) and [dl[k] for k in AR_MAPPING] == list(g["ar_delta"])
    random.seed(123)
    tot, per, cnt, _ = b200.compute_ar_vae_loss(z, attrs, AR_MAPPING, "subset", 40, AR_DG)
    assert abs(float(tot) - float(g["ars_total"])) <= TOL_LOSS * float(g["ars_total"])
    assert [cnt[k] for k in AR_MAPPING] == list(g["ars_cnt"])
    z4 = torch.from_numpy(g["ar_z4"]).to(DEV)
    tot4, _, _, _ = b200.compute_ar_vae_loss(z4, {k: v[:5] for k, v in attrs.items()}, AR_MAPPING, "all", None, AR_DG)
    assert abs(float(tot4) - float(g["ar_total4"])) <= TOL_LOSS * float(g["ar_total4"])
    # all attribute values equal -> no valid pair -> zero loss, zero count (reference semantics)
    t0, p0, c0, _ = b200.compute_ar_vae_loss(z, {"a0": torch.ones(12)}, {"a0": {"latent_channel": 0, "delta": 1.0}}, "all", None, None)
    assert float(t0) == 0.0 and c0["a0"] == 0


def test_ar_vae_loss_gradient_matches_oracle_autograd(b200, oracle):
    """The AR-VAE loss is a training regulariser (train_vae.py:407-415 adds ar_gamma * total to loss_g): its gradient
    w.r.t. the latents, through the device kernels, vs autograd through the oracle's restatement of losses.py:69-166
    (fp32 CPU).  "all" and seeded "subset" pairs, [B, C] and [B, C, H, W] latents, total and a per-attribute loss."""
    import random
    gen = torch.Generator().manual_seed(3)
    attrs = {"a0": torch.randint(20, 200, (12,), generator=gen).float(), "a1": torch.randint(20, 24, (12,), generator=gen).float(),
             "a2": torch.randint(20, 200, (12,), generator=gen).float()}
    for shape, mode, npairs in (((12, 10), "all", None), ((12, 10), "subset", 40), ((12, 10, 4, 4), "all", None)):
        z_cpu = torch.randn(shape, generator=gen).requires_grad_(True)
        z_gpu = z_cpu.detach().to(DEV).requires_grad_(True)
        random.seed(11)
        tot_r, per_r, _, _ = oracle.ar_vae_loss_ref(z_cpu, attrs, AR_MAPPING, mode, npairs, AR_DG)
        (tot_r + 0.5 * per_r["a2"]).backward()
        random.seed(11)
        tot, per, cnt, _ = b200.compute_ar_vae_loss(z_gpu, attrs, AR_MAPPING, mode, npairs, AR_DG)
        assert tot.requires_grad and per["a2"].requires_grad
        (tot + 0.5 * per["a2"]).backward()
        assert abs(float(tot.detach()) - float(tot_r.detach())) <= TOL_LOSS * float(tot_r.detach())
        g, gr = z_gpu.grad.cpu(), z_cpu.grad
        assert g.shape == gr.shape
        assert float((g - gr).abs().max()) <= 1e-5 + 1e-4 * float(gr.abs().max()), (shape, mode)
        untouched = [c for c in range(10) if c not in (0, 4, 9)]
        assert float(g[:, untouched].abs().max()) == 0.0
    # no gradient requested: plain tensors, no graph
    tot, _, _, _ = b200.compute_ar_vae_loss(z_gpu.detach(), attrs, AR_MAPPING, "all", None, AR_DG)
    assert not tot.requires_grad


def test_regressor_matches_reference_golden(b200, oracle):
    """LatentRegressor (row a18) vs outputs of the reference's own class; encode -> flatten -> head end to end vs the oracle."""
    g = np.load(GOLD / "regressor_ref.npz")
    x = torch.from_numpy(g["x"]).to(DEV)
    for tag, act, drop in (("relu", "relu", 0.1), ("gelu", "gelu", 0.0), ("lrelu", "leaky_relu", 0.0), ("elu", "elu", 0.0)):
        torch.manual_seed(77)
        reg = b200.LatentRegressor(4096, [256, 32], 6, dropout=drop, activation=act).to(DEV).eval()
        out = reg(x)
        ref = torch.from_numpy(g[tag])
        assert float((out.cpu() - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), tag   # fp32 path
    cfg = b200.config.AUTOENCODER_DEF_A
    ref_vae, vae = _models(b200, oracle, cfg)
    torch.manual_seed(3)
    head = b200.LatentRegressor(4 * 8 * 8, [64, 16], 6, dropout=0.1).to(DEV).eval()
    head_ref = oracle.LatentRegressorRef(4 * 8 * 8, [64, 16], 6, dropout=0.1).eval()
    head_ref.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    assert vae.encode_deterministic(torch.zeros(1, 1, 64, 64, device=DEV)).flatten(1).shape[1] == 256
    imgs = oracle.synthetic_images(3, 64, 64, seed=9)
    with torch.no_grad():
        want = head_ref(torch.flatten(ref_vae.encode(imgs)[0], 1))
    got = b200.regress_from_images(vae, head, imgs.to(DEV)).cpu()
    with pytest.raises(ValueError):
        b200.regress_from_images(vae, head, oracle.synthetic_images(1, 32, 32).to(DEV))
    assert float((got - want).norm() / want.norm()) <= TOL_LATENT


@pytest.mark.parametrize("hw", [384, 512])
def test_regression_sweep_extents(b200, oracle, hw):
    """BASELINE configs[4] (reg_edente_from_dente sweep, 256^2-512^2): deterministic encode -> flatten -> MLP and
    reconstruct_deterministic at the larger extents (attention L = 2304 / 4096), vs the oracle evaluated here."""
    cfg = b200.config.AUTOENCODER_DEF_A
    ref_vae, vae = _models(b200, oracle, cfg)
    lat = 4 * (hw // 8) ** 2
    torch.manual_seed(5)
    head = b200.LatentRegressor(lat, [256, 32], 6, dropout=0.1).to(DEV).eval()
    head_ref = oracle.LatentRegressorRef(lat, [256, 32], 6, dropout=0.1).eval()
    head_ref.load_state_dict({k: v.cpu() for k, v in head.state_dict().items()})
    x = oracle.synthetic_images(1, hw, hw, seed=13)
    with torch.no_grad():
        mu_r, _ = ref_vae.encode(x)
        want = head_ref(torch.flatten(mu_r, 1))
        rec_r = ref_vae.decode(mu_r)
    got = b200.regress_from_images(vae, head, x.to(DEV)).cpu()
    assert float((got - want).norm() / want.norm()) <= TOL_LATENT
    assert _rel_l2(vae.encode_deterministic(x.to(DEV)), mu_r) <= TOL_LATENT
    assert _rel_l2(vae.reconstruct_deterministic(x.to(DEV)), rec_r) <= TOL_RECON


def test_config_b_full_extent_and_ar_loss(b200, oracle):
    """BASELINE configs[3] (ar_vae_dente: latent 10, channels 64/128/256, attention L = 4096 d = 256) at 256^2:
    forward parity with the oracle and the AR loss on its z_mu."""
    cfg = b200.config.AUTOENCODER_DEF_B
    ref, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(2, 256, 256, seed=17)
    with torch.no_grad():
        mu_r, sg_r = ref.encode(x)
        eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(4))
        rec_r, _, _ = ref(x, eps)
    rec, mu, sg = vae.autoencoder(x.to(DEV), eps.to(DEV))
    assert mu.shape == (2, 10, 64, 64)
    assert _rel_l2(mu, mu_r) <= TOL_LATENT and _rel_l2(sg, sg_r) <= TOL_LATENT
    assert _rel_l2(rec, rec_r) <= TOL_RECON
    attrs = {"w": torch.tensor([31.0, 120.0]), "h": torch.tensor([77.0, 20.0])}
    mapping = {"w": {"latent_channel": 0, "delta": 1.0}, "h": {"latent_channel": 7, "delta": 2.0}}
    t_ref, _, c_ref, _ = oracle.ar_vae_loss_ref(mu_r, attrs, mapping, "all", None, None)
    t_got, _, c_got, _ = b200.compute_ar_vae_loss(mu, attrs, mapping, "all", None, None)
    assert c_got == c_ref and abs(float(t_got) - float(t_ref)) <= 5e-3 * max(1e-3, abs(float(t_ref)))


def test_eval_metrics_and_local_normalisation_match_reference_golden(b200, oracle):
    """fused PSNR/SSIM/MSE/MAE and the device LocalNormalizeByMask vs outputs of the
    reference's own functions (tests/golden/metrics_ref.npz)."""
    from pti_ldm_vae_b200 import eval_metrics, transforms
    g = np.load(GOLD / "metrics_ref.npz")
    for tag in "abc":
        img, rec = torch.from_numpy(g[f"{tag}_img"]).to(DEV), torch.from_numpy(g[f"{tag}_rec"]).to(DEV)
        m = eval_metrics.compute_eval_metrics(rec, img)
        assert np.allclose(m["psnr"].cpu().numpy(), g[f"{tag}_psnr"], atol=1e-3)
        assert np.allclose(m["ssim"].cpu().numpy(), g[f"{tag}_ssim"], atol=1e-4)
        assert np.allclose(m["mse"].cpu().numpy(), g[f"{tag}_mse"], rtol=1e-4)
        assert np.allclose(m["mae"].cpu().numpy(), g[f"{tag}_mae"], rtol=1e-4)
        assert np.allclose(eval_metrics.compute_psnr(rec, img, 2.0).cpu().numpy(), g[f"{tag}_psnr_raw"], atol=1e-3)
        assert np.allclose(eval_metrics.compute_ssim(rec, img, 2.0, 0.02, 0.05).cpu().numpy(), g[f"{tag}_ssim_raw"], atol=1e-4)
        m2 = eval_metrics.compute_eval_metrics(rec, img)
        assert all(torch.equal(m[k], m2[k]) for k in m), "deterministic"
    raw = torch.from_numpy(g["ln_raw"]).to(DEV)
    out = transforms.LocalNormalizeByMask()(raw)
    assert np.allclose(out.cpu().numpy(), g["ln_out"], rtol=1e-5, atol=2e-5)
    assert (out[:, :, :, :30] == 0).all() and float(out[3].abs().max()) == 0.0
    single = transforms.LocalNormalizeByMask()(raw[0, 0])
    assert torch.equal(single, out[0, 0])
    d = transforms.ApplyLocalNormd(["image"])({"image": raw.clone(), "other": 1})
    assert torch.equal(d["image"], out) and d["other"] == 1
    with pytest.raises(b200._lib.PtivaeError):
        eval_metrics.compute_psnr(rec.cpu(), img.cpu())


def test_pipelined_host_streaming_matches_serial(b200, oracle):
    """PipelinedVAE (H2D || replay || D2H over double buffers) returns, for every batch of a stream of different
    batches, exactly what the serial GraphedVAE call returns."""
    cfg = b200.config.AUTOENCODER_DEF_A
    _, vae = _models(b200, oracle, cfg)
    g = b200.GraphedVAE(vae, 2, 64, 64, mode="reconstruct")
    xs = [oracle.synthetic_images(2, 64, 64, seed=100 + i).pin_memory() for i in range(7)]
    want = [g(x.to(DEV)).clone().cpu() for x in xs]
    pipe = b200.PipelinedVAE(g)
    outs = [torch.empty(2, 1, 64, 64).pin_memory() for _ in xs]
    for x, o in zip(xs, outs):
        pipe.submit(x, o)
    pipe.synchronize()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
    with pytest.raises(ValueError):
        pipe.submit(torch.zeros(2, 1, 64, 64), outs[0])


def test_large_batch_is_bitwise_the_concatenation_of_small_batches(b200, oracle):
    """Size-independent property at more than the bench's full size: images are independent and every reduction has
    a fixed order, so a 160-image 256^2 batch (2.7 GB fp32 stream tensors: > 2^31 bytes, 32-bit offsets would wrap)
    must reproduce, bit for bit, what the same images give in batches of 64 + 64 + 32."""
    cfg = b200.config.AUTOENCODER_DEF_A
    _, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(160, 256, 256, seed=23).to(DEV)
    big = vae.reconstruct_deterministic(x)
    mu_big = vae.encode_deterministic(x)
    lo = 0
    for n in (64, 64, 32):
        part = vae.reconstruct_deterministic(x[lo:lo + n])
        assert torch.equal(part, big[lo:lo + n]), f"batch slice {lo}:{lo + n} differs"
        assert torch.equal(vae.encode_deterministic(x[lo:lo + n]), mu_big[lo:lo + n])
        lo += n
    assert torch.isfinite(big).all()


def test_module_cast_to_half_still_runs(b200, oracle):
    """A module cast with .half() (fp16 master parameters) still feeds the kernels correctly typed operands."""
    cfg = b200.config.AUTOENCODER_DEF_A
    ref, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(2, 64, 64, seed=5)
    with torch.no_grad():
        mu_r = ref.encode(x)[0]
    vae = vae.half()
    mu = vae.encode_deterministic(x.to(DEV))
    assert mu.dtype == torch.float32 and _rel_l2(mu, mu_r) <= 4 * TOL_LATENT      # masters themselves were rounded to fp16


def test_eval_metrics_multichannel_and_ragged_vs_oracle(b200, oracle):
    """The fused metrics kernel on shapes the reference's compute_ssim cannot take (C = 3; it is single-channel only),
    checked against the oracle restatement, plus extents that are not multiples of the 32-pixel tile."""
    from pti_ldm_vae_b200 import eval_metrics
    g = torch.Generator().manual_seed(31)
    for shape in ((2, 3, 45, 70), (1, 1, 31, 33), (3, 2, 64, 64)):
        img = torch.rand(*shape, generator=g)
        rec = (img + 0.07 * torch.randn(*shape, generator=g)).clamp(-0.2, 1.2)
        want = oracle.eval_metrics_ref(rec, img)
        got = eval_metrics.compute_eval_metrics(rec.to(DEV), img.to(DEV))
        assert np.allclose(got["ssim"].cpu().numpy(), want["ssim"].numpy(), atol=1e-4), shape
        assert np.allclose(got["psnr"].cpu().numpy(), want["psnr"].numpy(), atol=1e-3), shape
        assert np.allclose(got["mse"].cpu().numpy(), want["mse"].numpy(), rtol=1e-4)
        assert np.allclose(got["mae"].cpu().numpy(), want["mae"].numpy(), rtol=1e-4)


@pytest.mark.parametrize("dt,h,w,size", [(torch.uint8, 300, 517, (256, 256)), (torch.int16, 1024, 2048, (256, 256)),
                                         (torch.float32, 100, 90, (256, 256)), (torch.float32, 512, 512, (384, 384)),
                                         (torch.uint8, 77, 33, (20, 64))])
def test_resize_area_and_preprocess_batch(b200, oracle, dt, h, w, size):
    """the device Resize (MONAI Resize's default "area" mode delegates to
    torch.nn.functional.interpolate(mode="area"), which is the oracle here) for raw uint8 / uint16 / float32 batches,
    down- and up-scaling at non-integer ratios; and preprocess_batch = resize -> LocalNormalizeByMask (pinned to the
    reference transform by the golden test above) -> [B, 1, h, w]."""
    gen = torch.Generator().manual_seed(5)
    if dt == torch.float32:
        raw = torch.rand(3, h, w, generator=gen) * 1000.0
        as_float = raw
    elif dt == torch.uint8:
        raw = torch.randint(0, 256, (3, h, w), generator=gen).to(torch.uint8)
        as_float = raw.float()
    else:   # uint16 bits in an int16 tensor (values above 32767 included)
        vals = torch.randint(0, 65536, (3, h, w), generator=gen)
        raw = vals.to(torch.int32).numpy().astype("uint16").view("int16")
        raw = torch.from_numpy(raw)
        as_float = vals.float()
    as_float[:, :, : w // 4] = 0.0           # background band, as in the reference's data
    if dt != torch.float32:
        raw[:, :, : w // 4] = 0
    ref = F.interpolate(as_float[:, None], size=size, mode="area")[:, 0]
    out = b200.ops.resize_area(raw.to(DEV), size)
    assert out.shape == ref.shape and float((out.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    host = [r.numpy() for r in raw] if dt == torch.float32 else (raw.numpy().view("uint16") if dt == torch.int16 else raw.numpy())
    x = b200.transforms.preprocess_batch(host, size)
    assert x.shape == (3, 1) + tuple(size) and x.is_cuda and x.dtype == torch.float32
    expect = b200.transforms.LocalNormalizeByMask()(ref.to(DEV))
    assert float((x[:, 0] - expect).abs().max()) <= 1e-3
    vae = b200.VAEModel.from_config(b200.config.AUTOENCODER_DEF_A).to(DEV).eval()
    if size[0] % 8 == 0 and size[1] % 8 == 0:
        assert torch.isfinite(vae.encode_deterministic(x)).all()


def test_wide_levels_256_and_512(b200, oracle):
    """The constructor accepts channel widths up to 512; 256- and 512-wide levels take the unfused route
    (gn_apply -> conv_umma).  A small model with both widths against the oracle, forward and deterministic paths."""
    cfg = dict(spatial_dims=2, in_channels=1, out_channels=1, latent_channels=4, channels=[64, 256, 512],
               num_res_blocks=[1, 1, 1], norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, False],
               with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
    ref, vae = _models(b200, oracle, cfg)
    x = oracle.synthetic_images(2, 32, 48, seed=9)
    with torch.no_grad():
        mu_r, sg_r = ref.encode(x)
        eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(7))
        rec_r, _, _ = ref(x, eps)
    rec, mu, sg = vae.autoencoder(x.to(DEV), eps.to(DEV))
    assert _rel_l2(rec, rec_r) <= 1e-2 and _rel_l2(mu, mu_r) <= 5e-3 and _rel_l2(sg, sg_r) <= 5e-3


def test_regression_head_trains_over_the_frozen_encoder(b200, oracle):
    """reg_scripts' loop (regression_head.py:119-138 + an optimizer over the head): frozen kernel-backed encoder under
    no_grad, trainable MLP head with dropout.  In train mode the head is differentiable (ordinary autograd ops over the
    same nn.Sequential the inference kernel reads), the loss goes down, and eval-mode inference afterwards runs the kernel
    on the UPDATED parameters and agrees with the autograd evaluation."""
    cfg = b200.config.AUTOENCODER_DEF_A
    _, vae = _models(b200, oracle, cfg)
    for p in vae.parameters():
        p.requires_grad = False
    torch.manual_seed(5)
    head = b200.LatentRegressor(4 * 8 * 8, [64, 16], 6, dropout=0.1).to(DEV).train()
    opt = torch.optim.Adam(head.parameters(), lr=1e-2)
    x = oracle.synthetic_images(8, 64, 64, seed=8).to(DEV)
    target = torch.randn(8, 6, generator=torch.Generator().manual_seed(1)).to(DEV)
    losses = []
    for _ in range(30):
        opt.zero_grad(set_to_none=True)
        pred = b200.regress_from_images(vae, head, x)
        loss = torch.nn.functional.mse_loss(pred, target)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.5 * losses[0], losses
    head.eval()
    with torch.no_grad():
        out_kernel = b200.regress_from_images(vae, head, x)
        out_torch = head.mlp(torch.flatten(vae.encode_deterministic(x), start_dim=1))
    assert float((out_kernel - out_torch).abs().max()) <= 1e-4 * max(1.0, float(out_torch.abs().max()))
