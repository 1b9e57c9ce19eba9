"""CPU tests of the N>1 host logic with world_size 2 over gloo (no GPU): problem sharding of the batched mode,
max-over-ranks timing, result gather; and the partitioned mode's halo plan driven by real point-to-point messages between two
processes (what ncclSend / ncclRecv do on the GPUs) plus the all-reduce of the CG's dot products."""
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


def _partition_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesh_deform_b200 import meshgen as G, partition as PT
    from oracle.numpy_ref import cotan_weights
    P, F = G.grid_plane(31, 22)                              # every rank holds the global mesh on the host, like the engine's callers
    owner = PT.strip_owner(P, world)
    part = PT.build_local_part(F, owner, rank, world)
    x = np.random.default_rng(5).standard_normal((P.shape[0], 3))
    lx = np.full((part.n_local, 3), np.nan)
    lx[:part.n_owned] = x[part.owned_global]
    # halo exchange exactly as Transport::exchange: pack owned entries per neighbour, send; receive straight into the halo slots
    reqs, recv_bufs = [], []
    for k, q in enumerate(part.neighbor_rank):
        send = torch.from_numpy(np.ascontiguousarray(lx[part.send_index[part.send_offset[k]:part.send_offset[k + 1]]]))
        buf = torch.empty((int(part.recv_offset[k + 1] - part.recv_offset[k]), 3), dtype=torch.float64)
        recv_bufs.append(buf)
        reqs.append(dist.isend(send, int(q)))
        reqs.append(dist.irecv(buf, int(q)))
    for r in reqs:
        r.wait()
    for k, buf in enumerate(recv_bufs):
        lx[part.n_owned + part.recv_offset[k]:part.n_owned + part.recv_offset[k + 1]] = buf.numpy()
    Wl = cotan_weights(P[part.local_to_global], part.faces).tocsr()
    y = (Wl @ lx)[:part.n_owned]                              # owned rows of the one-ring operator
    partial = torch.tensor([(lx[:part.n_owned] * y).sum(), float(part.n_owned)], dtype=torch.float64)
    dist.all_reduce(partial)                                  # the CG's d.Ad, summed over the ranks
    np.savez(os.path.join(out_dir, "p%d.npz" % rank), gid=part.owned_global, y=y, dot=partial.numpy(), nan=np.isnan(lx).any())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partitioned_halo_exchange_over_gloo(tmp_path):
    from mesh_deform_b200 import meshgen as G
    from oracle.numpy_ref import cotan_weights
    world = 2
    mp.spawn(_partition_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    P, F = G.grid_plane(31, 22)
    x = np.random.default_rng(5).standard_normal((P.shape[0], 3))
    W = cotan_weights(P, F).tocsr()
    y_global = W @ x
    y = np.zeros_like(y_global)
    for r in range(world):
        z = np.load(tmp_path / ("p%d.npz" % r))
        assert not bool(z["nan"])                             # every halo slot was filled
        y[z["gid"]] = z["y"]
        assert abs(float(z["dot"][0]) - float((x * y_global).sum())) < 1e-9 * abs(float((x * y_global).sum()))
        assert int(z["dot"][1]) == P.shape[0]
    assert np.abs(y - y_global).max() < 1e-12
