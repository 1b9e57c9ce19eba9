"""CPU tests of the N>1 host logic with world_size 2 over gloo (no GPU): problem sharding of the batched mode,
max-over-ranks timing, result gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mesh_deform_b200.sharding import shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesh_deform_b200 import sharding
    begin, end = sharding.shard_range(n_items, rank, world)
    # stand-in for "deform my problems": result k is a deterministic function of the problem index
    local = np.stack([np.full((3,), float(k)) for k in range(begin, end)]) if end > begin else np.zeros((0, 3))
    local_ms = 10.0 + rank                                   # pretend device time
    worst = sharding.max_over_ranks(local_ms)
    total = sharding.sum_over_ranks(end - begin)
    gathered = sharding.gather_results(local, n_items)
    dist.barrier()
    if rank == 0:
        np.savez(os.path.join(out_dir, "r.npz"), worst=worst, total=total, gathered=gathered)
    dist.destroy_process_group()


def test_two_rank_batch_sharding_over_gloo(tmp_path):
    world, n_items = 2, 11
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    r = np.load(tmp_path / "r.npz")
    assert float(r["worst"]) == 11.0                          # max over ranks, not rank 0's own time
    assert float(r["total"]) == n_items
    assert np.array_equal(r["gathered"][:, 0], np.arange(n_items, dtype=float))
