"""Batch sharding for multi-GPU inference and the gradient all-reduce of data-parallel training: images are independent (GroupNorm and attention
are per-sample), so every rank owns a contiguous shard of the batch, weights are replicated and there is
NO data-path collective.  The only communication is host-side plumbing: a barrier and a max-reduce of the
per-rank elapsed time for benchmarking, and an optional gather of results.  Works with the ``nccl`` backend
on GPUs and with ``gloo`` on CPU (used by the tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of rank's shard of n items (first n % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_batch(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather of per-rank shards back into batch order (uneven shards are padded and trimmed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """Device-timed numbers are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# ---------------------------------------------------------------------------------------------------------------------
# data-parallel training (the reference: DistributedDataParallel around the autoencoder, train_vae.py:282; gradients are
# averaged over ranks by bucketed NCCL all-reduces).  Here the gradients of all parameters live in ONE flat fp32 buffer
# in ``parameters()`` order, so the reduction is a handful of contiguous segments.
# ---------------------------------------------------------------------------------------------------------------------
def gradient_segments(ae) -> dict:
    """[lo, hi) element ranges of the flat gradient buffer, grouped by when they become final during the backward:
    ``decoder`` (decoder stack + post_quant_conv: final once the decoder backward has finished, i.e. while the encoder
    backward still runs) and ``encoder`` (encoder stack + the two quant convs: final at the end)."""
    names = ["encoder", "decoder", "quant_conv_mu", "quant_conv_log_sigma", "post_quant_conv"]
    have = [n for n, _ in ae.named_children() if n in names]
    if have != names:
        raise RuntimeError(f"unexpected child order {have}: the flat-buffer segments assume {names}")
    order = [(n, getattr(ae, n)) for n in names]
    sizes = [sum(p.numel() for p in m.parameters()) for _, m in order]
    offs = [0]
    for sz in sizes:
        offs.append(offs[-1] + sz)
    return {"decoder": [(offs[1], offs[2]), (offs[4], offs[5])], "encoder": [(offs[0], offs[1]), (offs[2], offs[4])],
            "total": offs[5]}


def allreduce_segments(flat: torch.Tensor, segments, group=None) -> None:
    """Sum-all-reduce of the given [lo, hi) ranges of a flat buffer, in place (NCCL on GPUs, gloo in the CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for lo, hi in segments:
        if hi > lo:
            dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group)
