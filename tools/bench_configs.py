"""Forward throughput of the other BASELINE.json configurations (parity-test cases, not bench lines):
config A at the regression sweep's extents and config B (AR-VAE model) -- CUDA-graph replay, device-resident input."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import json
import torch
import _pkg
b200 = _pkg.load()
STREAM16 = "--fp32-stream" not in sys.argv      # inference default of bench.py: residual stream in the fp16 operand format
RECORDS = []
GFLOP = {("A", 256): 48.916, ("A", 384): 113.08, ("A", 512): 208.55, ("B", 256): 242.39}
for name, cfg, hw, b, mode in (("A", b200.config.AUTOENCODER_DEF_A, 256, 64, "forward"), ("A", b200.config.AUTOENCODER_DEF_A, 384, 32, "reconstruct"),
                               ("A", b200.config.AUTOENCODER_DEF_A, 512, 16, "reconstruct"), ("B", b200.config.AUTOENCODER_DEF_B, 256, 16, "forward"),
                               ("A", b200.config.AUTOENCODER_DEF_A, 256, 64, "encode")):
    torch.manual_seed(1234)
    vae = b200.VAEModel.from_config(cfg).cuda().eval()          # default-initialised weights (timing only)
    if STREAM16:
        vae.autoencoder.set_stream_dtype(torch.float16)
    g = b200.GraphedVAE(vae, b, hw, hw, mode=mode)
    g.x.copy_(torch.randn(b, 1, hw, hw, generator=torch.Generator().manual_seed(0)).cuda())
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gf = GFLOP[(name, hw)] * (17.46 / 48.916 if mode == "encode" else 1.0)
    print(f"config {name} {hw}x{hw} batch {b:3d} {mode:11s}: {ms:8.3f} ms/step  {b / ms * 1e3:9.1f} img/s  {gf * b / ms:7.1f} TFLOP/s (nominal)", flush=True)
    RECORDS.append({"config": name, "size": hw, "batch": b, "mode": mode, "stream": "fp16" if STREAM16 else "fp32", "ms_per_step": ms,
                    "images_per_s": b / ms * 1e3, "nominal_tflops": gf * b / ms, "launch": "CUDA graph replay, device-resident input, one B200"})
(pathlib.Path(__file__).resolve().parent.parent / "gpurun_out").mkdir(exist_ok=True)
(pathlib.Path(__file__).resolve().parent.parent / "gpurun_out" / "r2_configs.json").write_text(json.dumps(RECORDS, indent=1))
