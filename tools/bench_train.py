#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[2] core: vae_dente_2.json generator step without the stock-PyTorch
perceptual / adversarial nets): forward -> L1 + kl_weight*KL -> backward -> gradient all-reduce -> Adam, on N GPUs.

  python tools/bench_train.py [--batch 8] [--size 256] [--steps 20] [--warmup 3] [--config A|B] [--no-graph]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

One JSON line (rank 0): images/s (whole job, CUDA-graph replay of the whole step, CUDA-event timing, max over ranks),
the fwd / bwd / optimizer split from an eager pass with events around the phases, the per-op breakdown, and -- at N > 1
-- the step time with the all-reduce removed (what the exposed communication costs).
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GFLOP_FWD = {"A": 48.916, "B": 242.39}   # per image @256^2; training = 3x (row a20)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step (vae_dente_2.json: 8)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="A", choices=["A", "B"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="per-op CUDA-event breakdown of one eager step")
    ap.add_argument("--eager-baseline", action="store_true",
                    help="also time the stock-PyTorch (cuDNN) training step on this GPU in a subprocess (rank 0, 1 GPU)")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist
    import _pkg
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
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    cfg = b200.config.AUTOENCODER_DEF_A if args.config == "A" else b200.config.AUTOENCODER_DEF_B
    ref = aekl_ref.seeded_model(cfg, 1234)
    vae = b200.VAEModel.from_config(cfg)
    vae.load_state_dict(ref.state_dict(), strict=True)
    vae = vae.to(dev).train()
    B, S = args.batch, args.size
    x = b200.parallel.shard_batch(aekl_ref.synthetic_images(B * world, S, S, seed=0), rank, world).to(dev)
    kl_w = 1e-3 if args.config == "A" else 1e-4
    ts = b200.TrainStep(vae, lr=2.5e-5, kl_weight=kl_w, recon_loss="l1", overlap=not args.no_overlap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- eager phase split (events between the phases of one step) and optional per-op breakdown
    for _ in range(2):
        ts.step(x)
    barrier()
    phases = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    launches0 = b200.ops.LAUNCHES
    with torch.no_grad():
        run = b200.TrainRun(ts.ae)
        ev[0].record()
        recon, mu, sigma = run.forward(x)
        ev[1].record()
        rec_terms = b200.ops.l1l2(recon, x)
        kl = b200.ops.kl_loss(mu, sigma, True)
        d_recon = b200.ops.l1l2_bwd(recon, x, ts.gout_rec)
        d_mu, d_sigma = b200.ops.kl_bwd(mu, sigma, ts.gout_kl, True)
        ev[2].record()
        run.backward(d_recon, d_mu, d_sigma, ts.G, need_dx=False)
        ev[3].record()
        b200.ops.adam(ts.params, ts.grads, ts.m, ts.v, ts.step_dev, ts.lr, grad_scale=1.0 / world)
        ts.ae.refresh_packed()
        ev[4].record()
    torch.cuda.synchronize()
    launches_eager = b200.ops.LAUNCHES - launches0
    phases = {"forward_ms": ev[0].elapsed_time(ev[1]), "loss_ms": ev[1].elapsed_time(ev[2]),
              "backward_ms": ev[2].elapsed_time(ev[3]), "adam_ms": ev[3].elapsed_time(ev[4]),
              "note": "eager launches (host-bound at small batch); the graph replay below is the throughput number"}
    breakdown = None
    if args.breakdown and rank == 0:
        b200.ops.PROFILE = []
        ts.step(x)
        torch.cuda.synchronize()
        agg = {}
        for name, meta, e0, e1 in b200.ops.PROFILE:
            key = name if name not in ("wgrad", "conv_umma", "conv3x3_fused", "gn_bwd", "gn_apply") else f"{name}{meta}"
            a = agg.setdefault(key, [0.0, 0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
        b200.ops.PROFILE = None
        breakdown = {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]}
        tot = {}
        for k, v in agg.items():
            base = k.split("(")[0]
            t = tot.setdefault(base, [0.0, 0])
            t[0] += v[0]
            t[1] += v[1]
        breakdown = {"by_op": {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])},
                     "top_shapes": breakdown}

    # ---- throughput: CUDA-graph replay of the whole step (or eager with --no-graph)
    if not args.no_graph:
        ts.capture(B, S, S, warmup=1)
        stepfn = lambda: ts.replay()          # noqa: E731  (static input already holds this rank's shard)
        ts._static_x.copy_(x)
    else:
        stepfn = lambda: ts.step(x)           # noqa: E731
    for _ in range(max(args.warmup, 3)):
        out = stepfn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first_loss = float(out["recon_loss"])
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = stepfn()
    e1.record()
    barrier()
    ms = b200.parallel.max_over_ranks(e0.elapsed_time(e1), dev)
    last_loss = float(out["recon_loss"])
    if rank == 0:
        per_step = ms / args.steps
        value = B * world * args.steps / (ms * 1e-3)
        gf = GFLOP_FWD[args.config] * (S / 256.0) ** 2 * 3.0
        line = {"metric": "vae_train_step_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True,
                "scaling": "weak", "dtype": "f16 forward operands / bf16 backward operands, fp32 accumulate + master weights",
                "data": "synthetic",
                "config": {"workload": f"vae_dente_2.json-style generator step (config {args.config}): forward, L1 + {kl_w}*KL, backward, "
                                       f"gradient all-reduce, Adam; batch {B} per GPU, 1x{S}x{S}; LPIPS / PatchDiscriminator "
                                       "terms (stock PyTorch in the reference) not included",
                           "global_batch": B * world, "launch": "eager" if args.no_graph else "CUDA graph replay",
                           "all_reduce": "none (1 GPU)" if world == 1 else ("NCCL sum over the flat fp32 gradient buffer, decoder segment "
                                                                           "overlapped with the encoder backward" if ts.overlap else "NCCL, not overlapped")},
                "model_tflops_per_gpu": gf * 1e9 * value / world / 1e12,
                "phases_eager": phases, "launches_per_step_eager": launches_eager,
                "recon_loss_first_last": [first_loss, last_loss], "breakdown": breakdown}
        if args.eager_baseline and world == 1:
            import subprocess
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_eager_baseline.py"), "--mode", "train", "--batch", str(B),
                                "--size", str(S), "--config", args.config], capture_output=True, text=True, timeout=900)
            try:
                line["gpu_eager_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception:  # noqa: BLE001
                line["gpu_eager_baseline"] = {"error": (r.stderr or r.stdout)[-300:]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
