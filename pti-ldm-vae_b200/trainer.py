"""The VAE part of one training step of vae_scripts/train_vae.py (:380-445) as a fused, graph-capturable unit:

    recon, z_mu, z_sigma = autoencoder(images)            train_vae.py:385
    loss = L1|L2(recon, images) + kl_weight * compute_kl_loss(z_mu, z_sigma)      :393-394,419
    loss.backward()                                       :444   (+ DDP gradient all-reduce, :282)
    optimizer_g.step()                                    :445   (torch.optim.Adam, lr * world_size, :301)

``TrainStep`` owns ONE flat fp32 buffer each for parameters, gradients and the two Adam moments.  The backward
kernels write straight into the flat gradient buffer; the data-parallel reduction is one NCCL all-reduce per segment
of that buffer -- the decoder segment is issued on a side stream as soon as the decoder backward has finished, so
it travels over NVLink while the encoder backward runs; Adam is one kernel over the flat buffers.
The reference does the same work as 218 per-tensor autograd accumulations, DDP's bucketed reducer with a
``find_unused_parameters`` graph walk, and a per-tensor optimizer loop.

The perceptual / adversarial terms of the reference step (LPIPS-SqueezeNet, PatchDiscriminator) are outside the hot
path: a caller that needs them uses the autograd path (``AutoencoderKL.forward`` in train mode
returns differentiable tensors) and adds those losses with stock PyTorch.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops, parallel
from .training import FlatGrads, TrainRun


def flatten_parameters(module: torch.nn.Module) -> torch.Tensor:
    """Re-points every parameter of ``module`` at a slice of one flat fp32 buffer (values preserved) and returns it.
    ``state_dict()`` / ``load_state_dict()`` keep working: the parameters are ordinary views."""
    params = list(module.parameters())
    total = sum(p.numel() for p in params)
    flat = torch.empty(total, device=params[0].device, dtype=torch.float32)
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.detach().reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        off += n
    return flat


class TrainStep:
    """step(images) -> dict of device scalars (loss terms); parameters are updated in place."""

    def __init__(self, model, lr: float = 2.5e-5, kl_weight: float = 1e-3, recon_loss: str = "l1",
                 betas=(0.9, 0.999), eps: float = 1e-8, process_group=None, overlap: bool = True,
                 scale_lr_by_world: bool = True):
        ae = getattr(model, "autoencoder", model)
        self.ae = ae
        dev = next(ae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs the model on a CUDA device (there is no CPU path)")
        if recon_loss not in ("l1", "l2"):
            raise ValueError("recon_loss must be 'l1' or 'l2'")
        self.dev = dev
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.lr = lr * (self.world if scale_lr_by_world else 1)       # train_vae.py:301
        self.kl_weight, self.betas, self.eps = float(kl_weight), betas, eps
        self.overlap = overlap and self.world > 1
        self.params = flatten_parameters(ae)
        ae.invalidate_packed()
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.step_dev = torch.ones(1, device=dev, dtype=torch.float32)
        self.G = FlatGrads(ae, self.grads)
        # gradient-buffer split by the time the segments become final during the backward
        self.segments = parallel.gradient_segments(ae)
        assert self.segments["total"] == self.params.numel()
        self.comm = torch.cuda.Stream(device=dev) if self.overlap else None
        self.gout_rec = torch.tensor([1.0, 0.0] if recon_loss == "l1" else [0.0, 1.0], device=dev)
        self.gout_kl = torch.tensor([self.kl_weight], device=dev)
        self.recon_idx = 0 if recon_loss == "l1" else 1
        if ae._rng_dev is None:
            ae._rng_dev = torch.tensor([torch.initial_seed() & (2**63 - 1), 1], device=dev, dtype=torch.int64)
        if self.world > 1:
            # same starting point on every rank (DDP broadcasts at wrap time, train_vae.py:282)
            dist.broadcast(self.params, src=0, group=self.pg)
        self.graph = None
        self._static_x = None
        self._static_out = None

    # -- communication ------------------------------------------------------------------------------
    def _after_decoder(self) -> None:
        """Called by the backward once the decoder-side gradients are final: their all-reduce goes out on the
        communication stream and travels over NVLink while the encoder backward keeps the SMs busy."""
        if self.world == 1 or self.comm is None:
            return
        self.comm.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.comm):
            parallel.allreduce_segments(self.grads, self.segments["decoder"], self.pg)

    # -- one step -----------------------------------------------------------------------------------
    def _forward_backward(self, x: torch.Tensor, eps, after_decoder):
        """forward, loss terms, loss gradients, backward into the flat gradient buffer."""
        run = TrainRun(self.ae)
        recon, mu, sigma = run.forward(x, eps)
        xf = x.detach().contiguous().float()
        rec_terms = ops.l1l2(recon, xf)                     # (l1, l2)
        kl = ops.kl_loss(mu, sigma, True)
        d_recon = ops.l1l2_bwd(recon, xf, self.gout_rec)
        d_mu, d_sigma = ops.kl_bwd(mu, sigma, self.gout_kl, True)
        run.backward(d_recon, d_mu, d_sigma, self.G, need_dx=False, after_decoder=after_decoder)
        return {"recon_loss": rec_terms[self.recon_idx], "kl_loss": kl, "recon": recon, "z_mu": mu, "z_sigma": sigma}

    def _update(self) -> None:
        ops.adam(self.params, self.grads, self.m, self.v, self.step_dev, self.lr, self.betas, self.eps,
                 grad_scale=1.0 / self.world, advance=True)
        self.ae.refresh_packed()            # every weight pack from the updated masters: one launch

    def _reduce_encoder_and_join(self) -> None:
        if self.world == 1:
            return
        if self.comm is not None:
            parallel.allreduce_segments(self.grads, self.segments["encoder"], self.pg)
            torch.cuda.current_stream(self.dev).wait_stream(self.comm)
        else:
            parallel.allreduce_segments(self.grads, [(0, self.params.numel())], self.pg)

    def _step_impl(self, x: torch.Tensor, eps: torch.Tensor | None = None):
        with torch.no_grad():
            out = self._forward_backward(x, eps, self._after_decoder)
            self._reduce_encoder_and_join()
            self._update()
        return out

    def step(self, x: torch.Tensor, eps: torch.Tensor | None = None):
        return self._step_impl(x, eps)

    # -- CUDA graph ---------------------------------------------------------------------------------
    def capture(self, batch: int, height: int, width: int, warmup: int = 2):
        """Captures the step for a fixed shape; afterwards ``replay(x)`` copies x into the static input and replays.
        One GPU: ONE graph (forward, losses, backward, Adam, weight re-pack).  Several GPUs: THREE graphs -- forward +
        decoder backward | encoder backward | Adam + re-pack -- with the two NCCL all-reduces issued eagerly between
        them (the decoder segment on the communication stream, overlapping the second graph): NCCL calls stay outside
        the captures (capturing them dead-locked), at the price of two extra graph launches per step."""
        ae = self.ae
        self._static_x = torch.zeros((batch, ae.in_channels, height, width), device=self.dev, dtype=torch.float32)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self._static_x)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        if self.world == 1:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._static_out = self._step_impl(self._static_x)
            self._graphs = None
            return self
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=self.dev)
        cap.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(cap), torch.no_grad():
            g1.capture_begin()

            def split():                    # decoder-side gradients are final: close graph 1, open graph 2 (same pool)
                g1.capture_end()
                g2.capture_begin(pool=g1.pool())

            self._static_out = self._forward_backward(self._static_x, None, split)
            g2.capture_end()
            g3.capture_begin(pool=g1.pool())
            self._update()
            g3.capture_end()
        torch.cuda.current_stream(self.dev).wait_stream(cap)
        torch.cuda.synchronize(self.dev)
        self._graphs = (g1, g2, g3)
        self.graph = g1
        return self

    def replay(self, x: torch.Tensor | None = None):
        if x is not None:
            self._static_x.copy_(x, non_blocking=True)
        if getattr(self, "_graphs", None) is None:
            self.graph.replay()
            return self._static_out
        g1, g2, g3 = self._graphs
        g1.replay()
        self._after_decoder()               # all-reduce of the decoder segment on the communication stream
        g2.replay()                         # ... overlapped with the encoder backward
        self._reduce_encoder_and_join()
        g3.replay()
        return self._static_out
