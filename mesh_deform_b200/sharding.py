"""Host-side sharding helpers for multi-GPU runs (one process per GPU, torch.distributed for plumbing).

The hot path shards two ways:
  * batches of independent deformations (BASELINE.json configs[3]): contiguous ranges of problems per rank,
    topology replicated, NO data-path collective -- `shard_range`;
  * timing: every multi-GPU number is the max over ranks of a device-side time -- `max_over_ranks`.
Works with any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""


def shard_range(n_items, rank, world_size):
    """Contiguous [begin, end) of `n_items` problems owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, device=None):
    """MAX-all-reduce of a Python float over the default process group (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_results(local_array, n_items, device=None):
    """All-gather per-rank result blocks (numpy, first axis = local problems) into one (n_items, ...) array,
    in problem order. The only collective of the batched mode, after the solve."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_array
    world = dist.get_world_size()
    pieces = [None] * world
    dist.all_gather_object(pieces, local_array)
    out = np.concatenate(pieces, axis=0)
    assert out.shape[0] == n_items
    return out
