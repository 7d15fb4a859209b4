"""CUDA-graph replay of the whole encode -> sample -> decode pass for a fixed (B, H, W).

The eager path issues ~170 kernel launches + as many ctypes calls per forward; one graph launch
replaces them.  Sampling noise stays fresh across replays because the Philox
(seed, offset) pair lives in device memory and a 1-thread kernel bumps the offset inside the graph.
"""
from __future__ import annotations

import torch


class GraphedVAE:
    def __init__(self, model, batch: int, height: int, width: int, mode: str = "forward", warmup: int = 2):
        ae = getattr(model, "autoencoder", model)
        dev = next(ae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedVAE needs the model on a CUDA device")
        if mode not in ("forward", "reconstruct", "encode"):
            raise ValueError(mode)
        self.ae, self.mode = ae, mode
        self.x = torch.zeros((batch, ae.in_channels, height, width), device=dev, dtype=torch.float32)
        # ONE device-resident Philox (seed, offset) pair per model: every graph captured for this model bakes its address
        # into latent_sample / rng_advance, so it must never be replaced (a second GraphedVAE used to overwrite it and the
        # first graph then read / incremented freed memory); each graph also holds a reference.
        if ae._rng_dev is None or ae._rng_dev.device != dev:
            ae._rng_dev = torch.tensor([torch.initial_seed() & (2**63 - 1), 1], device=dev, dtype=torch.int64)
        self._rng = ae._rng_dev
        # packed weights are baked into the graph: remember the parameter versions to detect later in-place updates
        self._versions = tuple(p._version for p in ae.parameters())
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):  # packs weights, sets func attributes, warms the allocator
                    self._run()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()

    def _run(self):
        if self.mode == "forward":
            return self.ae(self.x)
        if self.mode == "reconstruct":
            return self.ae.reconstruct(self.x)
        return self.ae.encode(self.x)

    def __call__(self, x: torch.Tensor | None = None):
        """Copies x into the static input (if given), replays, returns the static outputs."""
        if tuple(p._version for p in self.ae.parameters()) != self._versions:
            raise RuntimeError("the model's parameters changed after this graph was captured (load_state_dict / optimizer "
                               "step): the graph replays the OLD packed weights -- capture a new GraphedVAE")
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out



class PipelinedVAE:
    """Streaming inference over host batches: H2D(i+1) || replay(i) || D2H(i-1).

    Wraps a GraphedVAE with two staging buffers on each side and two copy streams, so the PCIe transfers of
    neighbouring batches hide behind the kernels (the reference's inference_vae.py / evaluate_vae.py loop does
    `batch.to(device)` -> forward -> `.cpu()` serially).  Host tensors must be pinned.  `submit` only enqueues work;
    `synchronize` (or reading an output after its returned event) waits.
    """

    def __init__(self, graphed: GraphedVAE, output: int = 0):
        self.g, self.output = graphed, output
        dev = graphed.x.device
        out = graphed.out[output] if isinstance(graphed.out, (tuple, list)) else graphed.out
        self._out = out
        self.s_in = [torch.empty_like(graphed.x) for _ in range(2)]
        self.s_out = [torch.empty_like(out) for _ in range(2)]
        self.h2d, self.d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.compute = torch.cuda.current_stream(dev)
        mk = lambda: [torch.cuda.Event() for _ in range(2)]   # noqa: E731
        self.in_ready, self.in_free, self.out_ready, self.out_free = mk(), mk(), mk(), mk()
        self.i = 0

    def submit(self, x_host: torch.Tensor, out_host: torch.Tensor) -> torch.cuda.Event:
        """Enqueue one batch; returns the event after which out_host holds the result."""
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise ValueError("PipelinedVAE needs pinned host tensors")
        j = self.i & 1
        with torch.cuda.stream(self.h2d):
            if self.i >= 2:
                self.h2d.wait_event(self.in_free[j])          # the replay two steps ago has consumed this buffer
            self.s_in[j].copy_(x_host, non_blocking=True)
            self.in_ready[j].record(self.h2d)
        c = self.compute
        c.wait_event(self.in_ready[j])
        with torch.cuda.stream(c):
            self.g.x.copy_(self.s_in[j], non_blocking=True)   # device-to-device, microseconds
            self.in_free[j].record(c)
            self.g.graph.replay()
            if self.i >= 2:
                c.wait_event(self.out_free[j])                # the D2H two steps ago has drained this buffer
            self.s_out[j].copy_(self._out, non_blocking=True)
            self.out_ready[j].record(c)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(self.out_ready[j])
            out_host.copy_(self.s_out[j], non_blocking=True)
            self.out_free[j].record(self.d2h)
        self.i += 1
        return self.out_free[j]

    def synchronize(self) -> None:
        self.compute.wait_stream(self.h2d)
        self.compute.wait_stream(self.d2h)
        self.compute.synchronize()
