#!/usr/bin/env python
"""bench.py -- VAE encode -> sample -> decode images/sec on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full forward (AutoencoderKL.forward: encode, reparameterised sample, decode) over one
batch of 64 synthetic 1x256x256 images per GPU (configs[1]: vae_dente_no_adv, 16-bit tensor-core kernels).
Batches shard across ranks with no data-path collective (inference is independent per image): weak scaling.

JSON line (rank 0):
  value     images/s, whole job, inputs already resident in HBM, CUDA-graph replay, CUDA-event timing,
            max over ranks
  e2e       same metric through the public API with HOST (pinned) inputs: H2D copy of the batch and
            D2H copy of the reconstruction inside the timed region
  roofline  the dominant kernel class (largest share of the step): achieved algorithmic TFLOP/s or GB/s,
            measured live with CUDA events around every launch of one eager forward pass
  cpu_baseline  the CPU oracle (PyTorch fp32 restatement of the reference's MONAI model) on the host cores
--impl reference: the reference's CPU implementation of the path (the oracle port -- MONAI itself is not
installable offline) timed on the host cores with all threads, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "vae_encode_decode_images_per_sec"
UNIT = "images/s"
GFLOP_PER_IMG_A256 = 48.916  # SURVEY.md 8a row a1 (direct 9-tap count, config A @256^2)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_oracle_throughput(batch: int, size: int, threads: int, repeats: int, warmup: int):
    import torch
    from oracle import aekl_ref
    import _pkg
    cfg = _pkg.load().config.AUTOENCODER_DEF_A
    torch.set_num_threads(threads)
    model = aekl_ref.seeded_model(cfg, 1234)
    x = aekl_ref.synthetic_images(batch, size, size, seed=0)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            model(x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def run_reference(args) -> None:
    """--impl reference: the reference's CPU path (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sample_b = 8
    times = cpu_oracle_throughput(sample_b, args.size, threads, max(1, args.steps), max(1, min(args.warmup, 2)))
    total = sum(times)
    value = sample_b * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vae_dente_no_adv AutoencoderKL forward (encode->sample->decode), 1x%dx%d; CPU arm: bounded sample of "
                               "%d images per step of the same per-image workload (the b200 arm steps 64)" % (args.size, args.size, sample_b),
                   "batch_per_step": sample_b},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{len(times)} steps x {sample_b} images of the B=64 workload, torch {torch.__version__} CPU fp32, "
                                   "oracle/aekl_ref.py (MONAI 1.5.1 restatement; MONAI itself not installable offline)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- roofline
def survey_bytes(name, meta):
    """SURVEY.md 8(d) bytes of one launch: 16-bit activations, every tensor read once + written once per fused op
    (conv = (Cin*HW_in + Cout*HW_out) * 2 B; residuals / fp32 copies of this schedule are NOT counted)."""
    if name == "conv_umma":
        mode, n, h, w, cin, cout = meta
        ho, wo = (h // 2, w // 2) if mode == 1 else ((2 * h, 2 * w) if mode == 2 else (h, w))
        return 2.0 * n * (h * w * cin + ho * wo * cout)
    if name == "conv3x3_fused":
        n, h, w, cin, cout = meta[:5]
        return 2.0 * n * h * w * (cin + cout)
    if name == "conv3x3_fused_sc":
        n, h, w, cin, cout = meta[:5]
        return 2.0 * n * h * w * (cin + cout + meta[8])
    if name == "up2x_conv3x3":
        n, h, w, c, _ = meta
        return 2.0 * n * h * w * c * 5
    if name == "conv3x3_small_cin":
        n, h, w, cin, cout, _ = meta
        return n * h * w * (4.0 * cin + 2.0 * cout)
    if name == "conv3x3_small_cout":
        n, h, w, cin, cout, _ = meta
        return n * h * w * (2.0 * cin + 4.0 * cout)
    if name == "gn_stats":
        n, hw, c, _ = meta
        return 2.0 * n * hw * c
    if name == "gn_apply":
        n, hw, c, _, raw = meta
        return (4.0 + 2.0 * raw) * n * hw * c
    if name == "attention_fwd":
        b, l, d = meta
        return 8.0 * b * l * d
    return 0.0


def classify(name, meta):
    """-> (class key, algorithmic flops, algorithmic bytes of THIS schedule) for one launch."""
    if name == "conv_umma":
        mode, n, h, w, cin, cout = meta
        taps = {0: 9, 1: 9, 2: 9, 3: 1}[mode]
        ho, wo = (h // 2, w // 2) if mode == 1 else ((2 * h, 2 * w) if mode == 2 else (h, w))
        flops = 2.0 * n * ho * wo * cout * cin * taps          # nominal (direct-form) count, also for mode 2
        byt = 2.0 * n * (h * w * cin + ho * wo * cout) + 2.0 * taps * cin * cout
        return (f"conv{'3x3' if taps == 9 else '1x1'}_m{mode}_{cin}->{cout}@{h}x{w}", flops, byt)
    if name in ("gn_stats",):
        n, hw, c, esz = meta
        return (f"gn_stats_c{c}@{hw}", 0.0, float(esz) * n * hw * c)
    if name == "gn_apply":
        n, hw, c, esz, raw = meta
        return (f"gn_apply_c{c}@{hw}", 0.0, (esz + 2.0 + 2.0 * raw) * n * hw * c)
    if name == "conv3x3_fused":
        n, h, w, cin, cout, in_esz, out_esz, res_esz = meta
        flops = 2.0 * n * h * w * cout * cin * 9
        byt = n * h * w * (in_esz * cin + (out_esz + res_esz) * cout) + 2.0 * 9 * cin * cout
        return (f"fused3x3_{cin}->{cout}@{h}x{w}_in{in_esz}_res{res_esz}_out{out_esz}", flops, byt)
    if name == "conv3x3_fused_sc":
        n, h, w, cin, cout, in_esz, out_esz, _, sc = meta
        flops = 2.0 * n * h * w * cout * (cin * 9 + sc)
        byt = n * h * w * (in_esz * cin + 2.0 * sc + out_esz * cout) + 2.0 * cout * (9 * cin + sc)
        return (f"fused3x3+sc{sc}_{cin}->{cout}@{h}x{w}", flops, byt)
    if name == "up2x_conv3x3":
        n, h, w, c, e16 = meta
        out_b = 2.0 if e16 == 2 else (4.0 + 2.0 * e16)     # e16 == 2: 16-bit output only (16-bit residual stream)
        return (f"up2x_conv3x3_{c}@{h}x{w}_out{'2' if e16 == 2 else '4'}", 2.0 * n * 4 * h * w * c * c * 9,      # nominal direct-form count
                n * h * w * c * (2.0 + 4 * out_b) + 2.0 * 16 * c * c)
    if name == "conv3x3_small_cin":
        n, h, w, cin, cout, oesz = meta
        return (f"small_cin_{cin}->{cout}@{h}x{w}", 2.0 * n * h * w * cin * cout * 9, n * h * w * (4.0 * cin + float(oesz) * cout))
    if name == "conv3x3_small_cout":
        n, h, w, cin, cout, iesz = meta
        return (f"small_cout_{cin}->{cout}@{h}x{w}", 2.0 * n * h * w * cin * cout * 9, n * h * w * (float(iesz) * cin + 4.0 * cout))
    if name == "attention_fwd":
        b, l, d = meta
        return (f"attention_L{l}_d{d}", 4.0 * b * l * l * d, 8.0 * b * l * d)
    return (name, 0.0, 0.0)


def kernel_breakdown(model, x, passes: int):
    """One eager forward per pass with CUDA events around every launch -> per-class totals (ms)."""
    import torch
    import _pkg
    ops = _pkg.load().ops
    agg = {}
    for _ in range(passes):
        ops.PROFILE = []
        model(x)
        torch.cuda.synchronize()
        for name, meta, e0, e1 in ops.PROFILE:
            key, fl, by = classify(name, meta)
            a = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0, "survey_bytes": 0.0})   # sums over launches
            a["ms"] += e0.elapsed_time(e1)
            a["launches"] += 1
            a["flops"] += fl
            a["bytes"] += by
            a["survey_bytes"] += survey_bytes(name, meta)
        ops.PROFILE = None
    return agg


# --------------------------------------------------------------------------------------------- main arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    import _pkg
    from oracle import aekl_ref  # checker + cpu_baseline leg only

    b200 = _pkg.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator comes up; stdout carries exactly one
        # JSON line, so the banner goes to stderr (fd-level redirect: the banner is written by C code)
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

    B, S = args.batch, args.size
    cfg = b200.config.AUTOENCODER_DEF_A
    ref = aekl_ref.seeded_model(cfg, 1234)         # seeded random-init weights shared with the oracle
    vae = b200.VAEModel.from_config(cfg)
    vae.load_state_dict(ref.state_dict(), strict=True)
    vae = vae.to(dev).eval()
    if not args.fp32_stream:
        # inference: the residual stream between blocks is kept in the 16-bit operand format too (SURVEY 8d counts 16-bit
        # activations); the parity gate below runs with the same setting
        vae.autoencoder.set_stream_dtype(torch.float16)
    # each rank owns its contiguous shard of the global batch (weak scaling: B images per GPU)
    x_global = aekl_ref.synthetic_images(B * world, S, S, seed=0)
    x_host = b200.parallel.shard_batch(x_global, rank, world).clone().pin_memory()
    del x_global
    x_dev = x_host.to(dev)

    # parity gate on the benchmark's own weights (small sample, oracle on CPU)
    parity = None
    if rank == 0:
        xs = x_host[:1]
        with torch.no_grad():
            mu_r, sg_r = ref.encode(xs)
            eps = torch.randn(mu_r.shape, generator=torch.Generator().manual_seed(7))
            rec_r, _, _ = ref(xs, eps)
        rec, mu, sg = vae.autoencoder(xs.to(dev), eps.to(dev))
        rl = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())  # noqa: E731
        parity = {"recon_rel_l2": rl(rec, rec_r), "z_mu_rel_l2": rl(mu, mu_r), "z_sigma_rel_l2": rl(sg, sg_r)}

    launches_before = b200.ops.LAUNCHES
    graphed = b200.GraphedVAE(vae, B, S, S, mode="forward", warmup=2)
    launches_per_step = (b200.ops.LAUNCHES - launches_before) // 3   # 2 warm-ups + 1 capture
    graphed.x.copy_(x_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        graphed()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        graphed()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # ---- e2e: pinned host batch -> H2D -> forward -> reconstruction D2H, every step
    # (PipelinedVAE double-buffers both directions: every step still moves its own batch in and its own result out
    # inside the timed region, the transfers of neighbouring steps overlap the kernels)
    out_hosts = [torch.empty((B, 1, S, S), dtype=torch.float32).pin_memory() for _ in range(2)]
    pipe = b200.PipelinedVAE(graphed, output=0)
    for i in range(3):
        pipe.submit(x_host, out_hosts[i & 1])
    pipe.synchronize()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        pipe.submit(x_host, out_hosts[i & 1])
    pipe.synchronize()          # joins the copy streams into the current stream
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    with torch.no_grad():       # the last batch really arrived on the host
        e2e_check = float((out_hosts[(args.steps - 1) & 1].to(dev) - graphed.out[0]).abs().max())
    assert e2e_check == 0.0, f"pipelined host output differs from the device result ({e2e_check})"
    clocks = sampler.stop() if rank == 0 else None

    ms = b200.parallel.max_over_ranks(ms, dev)
    ms_e2e = b200.parallel.max_over_ranks(ms_e2e, dev)

    if rank == 0:
        total_images = B * world * args.steps
        value = total_images / (ms * 1e-3)
        e2e_value = total_images / (ms_e2e * 1e-3)
        # ---- roofline of the dominant kernel class, measured live (eager pass, events per launch)
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        vae.autoencoder._rng_dev = None
        agg = kernel_breakdown(vae.autoencoder, x_dev, passes=2)
        tot_ms = sum(a["ms"] for a in agg.values())
        top_key, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        tot_s = top["ms"] * 1e-3
        ai = top["flops"] / max(top["bytes"], 1.0)
        ridge = peaks.get("bf16_tflops_sustained", 1400.0) * 1e12 / (peaks.get("hbm_gbs", 6650.0) * 1e9)
        if top["flops"] > 0 and ai >= ridge:
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            roof = {"bound": "tensor", "achieved": top["flops"] / tot_s / 1e12, "peak": peak, "unit": "TFLOP/s"}
            roof["peak_source"] = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
        else:
            peak = peaks.get("hbm_gbs", 6650.0)
            roof = {"bound": "hbm", "achieved": top["bytes"] / tot_s / 1e9, "peak": peak, "unit": "GB/s"}
            roof["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
        roof["frac"] = roof["achieved"] / roof["peak"]
        # the same launches against SURVEY.md 8(d)'s per-kernel byte count (16-bit activations in and out, nothing else)
        roof["survey_8d"] = {"bytes_per_launch": top["survey_bytes"] / top["launches"],
                             "achieved_gbs": top["survey_bytes"] / tot_s / 1e9,
                             "frac_of_hbm_peak": top["survey_bytes"] / tot_s / 1e9 / peaks.get("hbm_gbs", 6650.0)}
        # DRAM traffic per launch of that kernel class from the committed ncu --set full capture, if any
        traffic_file = ROOT / "profiles" / "traffic.json"
        roof["traffic"] = None
        if traffic_file.exists():
            roof["traffic"] = json.loads(traffic_file.read_text()).get(top_key)
        roof["kernel"] = top_key
        roof["share_of_step"] = top["ms"] / tot_ms
        roof["avg_launch_ms"] = top["ms"] / top["launches"]
        roof["launches_per_step"] = top["launches"] // 2
        roof["algorithmic_bytes_per_launch"] = top["bytes"] / top["launches"]
        model_tflops = GFLOP_PER_IMG_A256 * (S / 256.0) ** 2 * 1e9 * value / world / 1e12 if S == 256 else None
        breakdown = {k: {"ms_per_step": v["ms"] / 2, "launches": v["launches"] // 2,
                         "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] else None,
                         "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9,
                         "survey_8d_gbs": v["survey_bytes"] / (v["ms"] * 1e-3) / 1e9}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        # whole-step view: sum over launches of max(bytes / HBM peak, flops / tensor peak) against the measured step
        hbm_pk, tc_pk = peaks.get("hbm_gbs", 6650.0) * 1e9, peaks.get("bf16_tflops_sustained", 1400.0) * 1e12
        floor_ms = sum(max(a["bytes"] / hbm_pk, a["flops"] / tc_pk) for a in agg.values()) / 2 * 1e3
        step_roofline = {"floor_ms": floor_ms, "frac": floor_ms / (ms / args.steps),
                         "algorithmic_gb_per_step": sum(a["bytes"] for a in agg.values()) / 2 / 1e9,
                         "algorithmic_tflop_per_step": sum(a["flops"] for a in agg.values()) / 2 / 1e12,
                         "note": "per-launch max(bytes/HBM peak, flops/tensor peak) summed over the step's launches "
                                 "(this schedule's own traffic) / measured step time"}
        out_dir = ROOT / "gpurun_out"
        out_dir.mkdir(exist_ok=True)
        (out_dir / "bench_breakdown.json").write_text(json.dumps({"eager_ms_per_step": tot_ms / 2, "classes": breakdown}, indent=1))

        cpu = None
        if world == 1 or True:
            threads = os.cpu_count() or 1
            times = cpu_oracle_throughput(4, S, threads, repeats=3, warmup=1)
            cpu = {"value": 4 / min(times), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"B=4 forward (configs[0]) best of 3 after 1 warm-up, oracle/aekl_ref.py fp32, torch CPU {threads} threads"}
        # ---- what a user would otherwise run on this GPU: the same op sequence in stock PyTorch (cuDNN / cuBLAS) --
        # own process, after every measurement of this one
        gpu_eager = None
        if world == 1 and not args.no_eager_baseline:
            try:
                r = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_eager_baseline.py"), "--mode", "infer",
                                    "--batch", str(B), "--size", str(S)], capture_output=True, text=True, timeout=600)
                gpu_eager = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as ex:  # noqa: BLE001
                gpu_eager = {"error": str(ex)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "vae_dente_no_adv.json AutoencoderKL forward (encode->sample->decode), 16-bit tensor-core "
                                   "kernels (fp16 operands: the bf16 operand format misses the 5e-3 z_mu tolerance, "
                                   "measured 9e-3; fp32 accumulate; residual stream "
                                   + ("fp32" if args.fp32_stream else "fp16 = the operand format") + "), "
                                   f"batch {B} per GPU, 1x{S}x{S}",
                       "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "no flush needed: every layer's activations (>=268 MB at 256^2) exceed the 126 MB L2",
                       "launch": "CUDA graph replay", "launches_per_step": launches_per_step},
            "clocks": clocks,
            "step_roofline": step_roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * S * S * 4, "d2h_bytes_per_step": B * S * S * 4,
                    "ms_per_step": ms_e2e / args.steps, "api": "PipelinedVAE(GraphedVAE(VAEModel)).submit(pinned host batch, pinned host recon): H2D(i+1) || kernels(i) || D2H(i-1); "
                           "the reconstruction is copied back every step, z_mu / z_sigma (2 x 1 MB per step) stay on the device"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager,
            "model_tflops_per_gpu": model_tflops,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer (default, BASELINE.json configs[1]: the contract line) | train (configs[2] core: forward + L1 + "
                         "KL + backward + gradient all-reduce + Adam, see tools/bench_train.py)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 64 infer / 8 train)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--fp32-stream", action="store_true",
                    help="keep the residual stream between blocks in fp32 (round-1 schedule) instead of fp16")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip the stock-PyTorch-on-the-same-GPU arm (a subprocess after the timed runs, N = 1 only)")
    args = ap.parse_args()
    if args.mode == "train" and args.impl == "b200":
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_train
        bench_train.main(["--batch", str(args.batch or 8), "--size", str(args.size), "--steps", str(args.steps),
                          "--warmup", str(args.warmup)])
        return
    if args.batch is None:
        args.batch = 64
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
