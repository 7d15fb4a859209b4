#!/usr/bin/env python
"""What a user of the reference would otherwise run on the same B200: the reference's op sequence (the oracle module =
MONAI 1.5.1 AutoencoderKL restated in stock PyTorch) on the GPU through cuDNN / cuBLAS / ATen,
  * fp32 eager, NCHW (exactly how the reference runs it: no autocast, no channels_last, no torch.compile), with TF32
    allowed for convolutions (the PyTorch default) and with TF32 off,
  * bf16 autocast + channels_last (a stronger, self-imposed bar: BASELINE.md section 1).
Inference: forward of B images; training: forward + L1 + KL + backward + torch.optim.Adam step.
Prints one JSON object.  Runs in its own process: tools/ and bench.py call it as a subprocess AFTER their own timing."""
from __future__ import annotations

import argparse
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--config", default="A")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F
    import _pkg
    from oracle import aekl_ref
    b200 = _pkg.load()
    cfg = b200.config.AUTOENCODER_DEF_A if args.config == "A" else b200.config.AUTOENCODER_DEF_B
    dev = torch.device("cuda:0")
    x = aekl_ref.synthetic_images(args.batch, args.size, args.size, seed=0).to(dev)
    out = {"mode": args.mode, "batch": args.batch, "size": args.size, "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "unit": "images/s"}

    def bench(name, tf32, autocast, channels_last):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True               # the reference sets it (train_vae.py:90)
        model = aekl_ref.seeded_model(cfg, 1234).to(dev)
        xin = x
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
            xin = x.contiguous(memory_format=torch.channels_last)
        opt = torch.optim.Adam(model.parameters(), lr=2.5e-5) if args.mode == "train" else None

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                if args.mode == "infer":
                    with torch.no_grad():
                        return model(xin)[0]
                opt.zero_grad(set_to_none=True)
                recon, mu, sigma = model(xin)
                loss = F.l1_loss(recon.float(), xin) + 1e-3 * aekl_ref.kl_loss_ref(mu.float(), sigma.float())
            loss.backward()
            opt.step()
            return loss

        try:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out[name] = {"images_per_s": args.batch / (ms * 1e-3), "ms_per_step": ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
        except Exception as ex:  # noqa: BLE001  (e.g. out of memory at a large batch)
            out[name] = {"error": str(ex)[:200]}
        del model, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()

    bench("fp32_eager_tf32_convs", True, False, False)
    bench("fp32_eager_no_tf32", False, False, False)
    bench("bf16_autocast_channels_last", True, True, True)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
