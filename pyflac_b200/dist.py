"""Multi-GPU plumbing: streams are independent, so ranks shard them with NO data-path collective;
torch.distributed is used only for the barrier and the max-over-ranks reduction of timings / counters."""


def shard_range(n_items, rank, world):
    """Contiguous block of items for `rank` (keeps a stream's frames, MD5 and callback order on one GPU)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_max(values, device=None):
    """Element-wise max over ranks of a list of floats (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def reduce_sum(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]
