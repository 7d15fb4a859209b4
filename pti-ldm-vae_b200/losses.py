"""Loss entry points with the reference's names and semantics
(/root/reference/src/pti_ldm_vae/models/losses.py:4-66; L1/MSE: vae_scripts/train_vae.py:289-296).
The reductions run as deterministic two-stage CUDA kernels (csrc/latent_loss.cu)."""
from __future__ import annotations

import torch

from . import ops


class _KLFn(torch.autograd.Function):
    """compute_kl_loss with its gradient kernel; the upstream gradient stays on the device (no host sync)."""

    @staticmethod
    def forward(ctx, mu, t, input_is_logvar):
        mu_c, t_c = mu.detach().contiguous().float(), t.detach().contiguous().float()
        ctx.save_for_backward(mu_c, t_c)
        ctx.is_logvar = bool(input_is_logvar)
        return ops.kl_loss(mu_c, t_c, ctx.is_logvar)

    @staticmethod
    def backward(ctx, g):
        mu, t = ctx.saved_tensors
        dmu, dt = ops.kl_bwd(mu, t, g.detach().reshape(1).float().contiguous(), ctx.is_logvar)
        return dmu, dt, None


class _L1L2Fn(torch.autograd.Function):
    """(mean |a-b|, mean (a-b)^2) as one pass; gradient w.r.t. both arguments."""

    @staticmethod
    def forward(ctx, a, b):
        a_c, b_c = a.detach().contiguous().float(), b.detach().contiguous().float()
        ctx.save_for_backward(a_c, b_c)
        return ops.l1l2(a_c, b_c)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        d = ops.l1l2_bwd(a, b, g.detach().float().contiguous())
        return (d if ctx.needs_input_grad[0] else None), (-d if ctx.needs_input_grad[1] else None)


class _SpatialMeanFn(torch.autograd.Function):
    """latent_vectors = z_mu.mean((2, 3)) with its (broadcast) gradient kernel."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return ops.spatial_mean(x)

    @staticmethod
    def backward(ctx, g):
        return ops.spatial_mean_bwd(g, ctx.shape)


class _ARFn(torch.autograd.Function):
    """All attributes of the AR-VAE loss in one kernel, differentiable w.r.t. the latent vectors [B, C].
    Returns (total [scalar], per-attribute losses [L], pair counts int32 [L])."""

    @staticmethod
    def forward(ctx, zbar, attrs, ch, dl, pairs):
        z = zbar.detach().contiguous().float()
        loss, cnt, tot = ops.ar_vae_loss(z, attrs, ch, dl, pairs)
        ctx.save_for_backward(z, attrs, ch, dl, cnt)
        ctx.pairs = pairs
        ctx.mark_non_differentiable(cnt)
        return tot[0], loss, cnt

    @staticmethod
    def backward(ctx, g_total, g_attr, _g_cnt):
        z, attrs, ch, dl, cnt = ctx.saved_tensors
        gt = None if g_total is None else g_total.detach().reshape(1).float().contiguous()
        ga = None if g_attr is None else g_attr.detach().float().contiguous()
        if gt is None and ga is None:
            return None, None, None, None, None
        return ops.ar_vae_loss_bwd(z, attrs, ch, dl, ctx.pairs, cnt, gt, ga), None, None, None, None


def _wants_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


def compute_kl_loss(z_mu: torch.Tensor, z_logvar: torch.Tensor, *, input_is_logvar: bool = True) -> torch.Tensor:
    """KL of a diagonal Gaussian, batch mean.  As the reference calls it (train_vae.py:394) the second
    argument is sigma interpreted as log-variance; that quirk is preserved bit-for-formula.  Differentiable."""
    if _wants_grad(z_mu, z_logvar):
        return _KLFn.apply(z_mu, z_logvar, input_is_logvar)
    return ops.kl_loss(z_mu, z_logvar, input_is_logvar)


def l1_and_mse(a: torch.Tensor, b: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    r = _L1L2Fn.apply(a, b) if _wants_grad(a, b) else ops.l1l2(a, b)
    return r[0], r[1]


def l1_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.L1Loss() (mean), differentiable."""
    return l1_and_mse(a, b)[0]


def mse_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.MSELoss() (mean), differentiable."""
    return l1_and_mse(a, b)[1]


def compute_total_loss(recons_loss, kl_loss, perceptual_loss, adv_gen_loss, ar_loss, *, kl_weight: float,
                       perceptual_weight: float, adv_weight: float, ar_gamma: float, ar_vae_enabled: bool):
    """losses.py:33-66 -- scalar arithmetic, stays in PyTorch."""
    total = recons_loss + kl_weight * kl_loss + perceptual_weight * perceptual_loss + adv_weight * adv_gen_loss
    if ar_vae_enabled:
        total = total + ar_gamma * ar_loss
    return total


def compute_ar_vae_loss(latent_vectors: torch.Tensor, attributes: dict, attribute_latent_mapping: dict, pairwise_mode: str,
                        subset_pairs, delta_global):
    """Attribute-regularised VAE loss with the reference's signature, validation and return tuple
    (/root/reference/src/pti_ldm_vae/models/losses.py:69-166): ``(total_loss, losses_per_attr, pair_counts, deltas_per_attr)``.
    All attributes are evaluated by ONE kernel (no host pair lists in "all" mode, one device->host read for the
    pair counts instead of two syncs per attribute).  "subset" mode samples pairs on the host with
    ``random.sample`` exactly like the reference, so a seeded ``random`` gives the same pairs."""
    grad = _wants_grad(latent_vectors)      # the loss is a training regulariser (train_vae.py:407-415): keep the graph
    if latent_vectors.dim() == 4:
        latent_vectors = _SpatialMeanFn.apply(latent_vectors.contiguous().float()) if grad else ops.spatial_mean(latent_vectors)
    elif latent_vectors.dim() != 2:
        raise ValueError(f"Expected latent shape [B, C] or [B, C, H, W], got {latent_vectors.shape}")
    latent_vectors = latent_vectors.contiguous().float() if grad else latent_vectors.detach().contiguous().float()
    batch_size, latent_dim = latent_vectors.shape
    if pairwise_mode not in {"all", "subset"}:
        raise ValueError(f"pairwise must be 'all' or 'subset', got {pairwise_mode}")
    if pairwise_mode == "subset" and (subset_pairs is None or subset_pairs <= 0):
        raise ValueError("subset_pairs must be a positive integer when pairwise='subset'")
    dev = latent_vectors.device
    names, chans, deltas, rows = [], [], [], []
    for attr_name, mapping in attribute_latent_mapping.items():
        target_latent = int(mapping["latent_channel"])
        if target_latent >= latent_dim:
            raise ValueError(f"Latent channel {target_latent} for attribute {attr_name} exceeds latent size {latent_dim}")
        attr_values = attributes.get(attr_name)
        if attr_values is None:
            raise KeyError(f"Missing attribute values for {attr_name} in batch.")
        delta_attr = mapping.get("delta")
        if delta_attr is None and delta_global and delta_global.get("enabled", False):
            delta_attr = delta_global.get("value")
        if delta_attr is None:
            raise ValueError(f"Delta not provided for {attr_name} and no delta_global fallback.")
        names.append(attr_name)
        chans.append(target_latent)
        deltas.append(float(delta_attr))
        rows.append(torch.as_tensor(attr_values, dtype=torch.float32).to(dev).reshape(-1))
    if not names:
        return torch.tensor(0.0, device=dev), {}, {}, {}
    zero = torch.tensor(0.0, device=dev)
    if batch_size < 2:
        return zero, {n: zero.clone() for n in names}, {n: 0 for n in names}, dict(zip(names, deltas))
    attrs = torch.stack(rows).contiguous()
    ch = torch.tensor(chans, device=dev, dtype=torch.int32)
    dl = torch.tensor(deltas, device=dev, dtype=torch.float32)
    if pairwise_mode == "all":
        total, per, cnt = _ARFn.apply(latent_vectors, attrs, ch, dl, None)
        cnt_h = cnt.tolist()
    else:
        import random
        all_pairs = [(i, j) for i in range(batch_size) for j in range(batch_size) if i != j]
        per_list, cnt_h = [], []
        for k in range(len(names)):           # the reference draws a fresh sample per attribute, in mapping order
            pairs = random.sample(all_pairs, min(len(all_pairs), int(subset_pairs)))
            pt = torch.tensor(pairs, device=dev, dtype=torch.int32).contiguous()
            _, l1, c1 = _ARFn.apply(latent_vectors, attrs[k:k + 1].contiguous(), ch[k:k + 1].contiguous(),
                                    dl[k:k + 1].contiguous(), pt)
            per_list.append(l1[0])
            cnt_h.append(int(c1[0]))
        per = torch.stack(per_list)
        total = per.sum()
    return (total, {n: per[k] for k, n in enumerate(names)}, {n: int(cnt_h[k]) for k, n in enumerate(names)},
            dict(zip(names, deltas)))
