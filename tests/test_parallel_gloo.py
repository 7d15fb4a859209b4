"""N>1 host logic on CPU: world_size-2 gloo.  Ranks shard a batch with no data-path collective, results
gather back in order, timing reduces with MAX (bench.py contract)."""
import os
import pathlib
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, str(ROOT))
    import _pkg
    par = _pkg.load().parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(n_total * 6, dtype=torch.float32).view(n_total, 1, 2, 3)
        mine = par.shard_batch(x)
        lo, hi = par.shard_bounds(n_total, rank, world)
        assert mine.shape[0] == hi - lo and torch.equal(mine, x[lo:hi])
        # stand-in for the per-shard encode->decode (independent per image): any per-sample map
        local = mine * 2 + 1
        full = par.gather_batch(local, n_total)
        assert torch.equal(full, x * 2 + 1)
        t = par.max_over_ranks(10.0 + rank)
        assert t == 10.0 + world - 1
        dist.barrier()
        (pathlib.Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_two_rank_sharding(tmp_path, n_total):
    port = 29500 + (os.getpid() % 400) + n_total
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, str(ROOT))
    import _pkg
    par = _pkg.load().parallel
    for n in (0, 1, 5, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [par.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        par.shard_bounds(4, 2, 2)


# ------------------------------------------------------------------------------------------- training: gradient all-reduce
def _grad_worker(rank, world, port, out_dir):
    """Data-parallel gradient reduction on the flat buffer (the host logic of TrainStep, on CPU with gloo): the segment
    all-reduces cover every parameter exactly once, and sum / world equals the single-process gradient of the
    concatenated batch (what DDP's mean all-reduce gives the reference, train_vae.py:282,444)."""
    sys.path.insert(0, str(ROOT))
    import _pkg
    from oracle import aekl_ref
    b200 = _pkg.load()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = dict(b200.config.AUTOENCODER_DEF_A)
        cfg.update(channels=[32, 64], attention_levels=[False, False], num_res_blocks=1)
        ref = aekl_ref.seeded_model(cfg, 1234)                      # CPU stand-in for the kernels (same parameter tree)
        holder = b200.AutoencoderKL(**{k: v for k, v in cfg.items() if not k.startswith("_")})   # parameter holder only
        holder.load_state_dict(ref.state_dict(), strict=True)
        seg = b200.parallel.gradient_segments(holder)
        covered = sorted(seg["decoder"] + seg["encoder"])
        assert covered[0][0] == 0 and covered[-1][1] == seg["total"] == sum(p.numel() for p in holder.parameters())
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:])), covered
        x = aekl_ref.synthetic_images(4, 32, 32, seed=0)
        eps = torch.randn(4, cfg["latent_channels"], 16, 16, generator=torch.Generator().manual_seed(7))

        def grads_of(xs, es):
            ref.zero_grad(set_to_none=True)
            recon, mu, sigma = ref(xs, es)
            (torch.nn.functional.l1_loss(recon, xs) + 1e-3 * aekl_ref.kl_loss_ref(mu, sigma)).backward()
            named = dict(ref.named_parameters())
            return torch.cat([named[n].grad.reshape(-1) for n, _ in holder.named_parameters()])

        lo, hi = b200.parallel.shard_bounds(4, rank, world)
        flat = grads_of(x[lo:hi], eps[lo:hi]).clone()
        G = b200.FlatGrads(holder, flat)                            # views in parameters() order over the same buffer
        assert G.flat.data_ptr() == flat.data_ptr() and len(G.order) == len(list(holder.parameters()))
        b200.parallel.allreduce_segments(flat, seg["decoder"])      # what TrainStep issues after the decoder backward
        b200.parallel.allreduce_segments(flat, seg["encoder"])      # ... and at the end
        full = grads_of(x, eps)
        err = float((flat / world - full).norm() / full.norm())
        assert err < 1e-5, err
        (pathlib.Path(out_dir) / f"gok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "gok0").exists() and (tmp_path / "gok1").exists()
