#!/usr/bin/env python
"""BASELINE.json configs[2] in full: the vae_dente_2.json generator + discriminator step of vae_scripts/train_vae.py:380-458
(L1 + 0.001*KL + 1.0*LPIPS-squeeze + 3.0*LSGAN generator loss; then the PatchDiscriminator step), batch 8 per GPU, 1x256x256.

  python tools/bench_train_full.py [--batch 8] [--steps 10] [--arm b200|torch|both]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train_full.py ...

Two arms on the same GPU(s), same auxiliary networks (tools/aux_nets.py: stock PyTorch, random init -- SURVEY.md 8d C3):
  b200  : the VAE is this repo's kernels behind the reference's loop -- `recon, mu, sigma = vae(x)` in train mode is the
          autograd edge (VAEFunction), losses via b200.l1_loss / compute_kl_loss, `loss_g.backward()`, torch.optim.Adam;
          multi-GPU: torch DistributedDataParallel(find_unused_parameters=True) around both nets, as train_vae.py:279-282
  torch : the VAE is the oracle module (stock PyTorch / cuDNN, TF32 convs as cudnn.benchmark picks them)
One JSON line: images/s per arm, and for the b200 arm the split VAE forward / aux forward / backward (aux + VAE) /
optimizers / discriminator step from CUDA events of an eager step.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--arm", default="both", choices=["b200", "torch", "both"])
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP
    import _pkg
    import aux_nets
    from oracle import aekl_ref

    b200 = _pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.backends.cudnn.benchmark = True          # train_vae.py:90
    cfg = b200.config.AUTOENCODER_DEF_A
    B, S = args.batch, args.size
    x = b200.parallel.shard_batch(aekl_ref.synthetic_images(B * world, S, S, seed=0), rank, world).to(dev)
    KL_W, P_W, ADV_W, LR = 1e-3, 1.0, 3.0, 2.5e-5 * world

    def build(arm):
        torch.manual_seed(1234)
        ref = aekl_ref.seeded_model(cfg, 1234)
        if arm == "b200":
            vae = b200.VAEModel.from_config(cfg)
            vae.load_state_dict(ref.state_dict(), strict=True)
            vae = vae.to(dev).train()
        else:
            vae = ref.to(dev).train()
        disc = aux_nets.PatchDiscriminatorRef().to(dev).train()
        lp = aux_nets.LPIPSSqueezeRef().to(dev).eval()
        if world > 1:
            vae = DDP(vae, device_ids=[local], output_device=local, find_unused_parameters=True)
            disc = DDP(disc, device_ids=[local], output_device=local, find_unused_parameters=True)
        og = torch.optim.Adam(vae.parameters(), lr=LR)
        od = torch.optim.Adam(disc.parameters(), lr=LR)
        return vae, disc, lp, og, od

    def make_step(arm, vae, disc, lp, og, od, ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()

        def step():
            og.zero_grad(set_to_none=True)
            mark(0)
            recon, mu, sigma = vae(x)
            mark(1)
            if arm == "b200":
                rec = b200.l1_loss(recon, x)
                kl = b200.compute_kl_loss(mu, sigma)
            else:
                rec = F.l1_loss(recon, x)
                kl = aekl_ref.kl_loss_ref(mu, sigma)
            p = lp(aux_nets.ensure_three_channels(recon.float()), aux_nets.ensure_three_channels(x))
            g = aux_nets.lsgan(disc(recon.contiguous().float())[-1], True)
            loss_g = b200.compute_total_loss(rec, kl, p, g, 0.0, kl_weight=KL_W, perceptual_weight=P_W, adv_weight=ADV_W,
                                             ar_gamma=0.0, ar_vae_enabled=False)
            mark(2)
            loss_g.backward()
            mark(3)
            og.step()
            mark(4)
            od.zero_grad(set_to_none=True)
            lf = aux_nets.lsgan(disc(recon.contiguous().detach())[-1], False)
            lr_ = aux_nets.lsgan(disc(x.contiguous())[-1], True)
            (ADV_W * 0.5 * (lf + lr_)).backward()
            od.step()
            mark(5)
            return loss_g
        return step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"metric": "vae_full_train_step_images_per_sec", "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "config": {"workload": f"vae_dente_2.json generator + discriminator step (train_vae.py:380-458), config A, batch {B} per GPU, "
                                  f"1x{S}x{S}; LPIPS-squeeze and PatchDiscriminator: stock PyTorch, random init (tools/aux_nets.py)",
                      "global_batch": B * world, "launch": "eager (reference loop)"},
           "data": "synthetic", "scaling": "weak"}
    for arm in (["b200", "torch"] if args.arm == "both" else [args.arm]):
        nets = build(arm)
        step = make_step(arm, *nets)
        for _ in range(max(args.warmup, 3)):
            loss = step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        barrier()
        ms = b200.parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
        rec = {"images_per_s": B * world / (ms * 1e-3), "ms_per_step": ms, "loss_g": float(loss.detach()),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        make_step(arm, *nets, ev=ev)()
        torch.cuda.synchronize()
        rec["split_ms"] = {"vae_forward": ev[0].elapsed_time(ev[1]), "losses_and_aux_forward": ev[1].elapsed_time(ev[2]),
                           "backward_aux_plus_vae": ev[2].elapsed_time(ev[3]), "optimizer_g": ev[3].elapsed_time(ev[4]),
                           "discriminator_step": ev[4].elapsed_time(ev[5])}
        out[arm] = rec
        del nets, step
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if rank == 0:
        if "b200" in out:
            out["value"] = out["b200"]["images_per_s"]
            out["ms_per_step"] = out["b200"]["ms_per_step"]
        if "b200" in out and "torch" in out:
            out["speedup_vs_stock_pytorch_same_gpu"] = out["b200"]["images_per_s"] / out["torch"]["images_per_s"]
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
