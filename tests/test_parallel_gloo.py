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
