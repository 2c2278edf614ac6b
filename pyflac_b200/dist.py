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


# ---------------------------------------------------------------------------------------------------------------
# Scatter / gather over NCCL (NVLink 5 / NVSwitch) for PCM that is born on ONE GPU: the root deals
# contiguous blocks of streams to the ranks, every rank encodes its block, the packed bytes come back to the root.
# The codec itself has no exchange step; when every rank loads its own streams none of this is needed.
def scatter_streams(pcm_root, n_streams, elems_per_stream, dtype, device, root=0):
    """Root holds `pcm_root`: tensor (n_streams, elems_per_stream) on `device`.  Returns this rank's block
    (tensor (k, elems_per_stream)) and its (lo, hi) stream range.  Works on NCCL (device tensors) and gloo (CPU)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    spans = [shard_range(n_streams, r, world) for r in range(world)]
    kmax = max(hi - lo for lo, hi in spans)
    out = torch.empty((kmax, elems_per_stream), dtype=dtype, device=device)
    out_b = out.view(torch.uint8)                   # bytes on the wire: every backend moves uint8, not every one moves int16
    if rank == root:
        parts = []
        for lo, hi in spans:
            blk = pcm_root[lo:hi]
            if hi - lo < kmax:                      # dist.scatter wants equal shapes: pad the short blocks
                pad = torch.zeros((kmax - (hi - lo), elems_per_stream), dtype=dtype, device=device)
                blk = torch.cat([blk, pad], 0)
            parts.append(blk.contiguous().view(torch.uint8))
        dist.scatter(out_b, scatter_list=parts, src=root)
    else:
        dist.scatter(out_b, scatter_list=None, src=root)
    lo, hi = spans[rank]
    return out[:hi - lo], (lo, hi)


def gather_packed(arena, nbytes, device, root=0):
    """Every rank contributes the first `nbytes` bytes of its uint8 tensor `arena`; the root gets (list of uint8 tensors
    in rank order, sizes), the others (None, sizes).  Sizes travel with an all_gather, payloads with point-to-point
    sends, so nothing is padded to the largest rank."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.tensor([int(nbytes)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [int(s.item()) for s in sizes]
    if rank == root:
        bufs = [arena[:nbytes] if r == root else torch.empty(sizes[r], dtype=torch.uint8, device=device) for r in range(world)]
        ops = [dist.P2POp(dist.irecv, bufs[r], r) for r in range(world) if r != root and sizes[r] > 0]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return bufs, sizes
    if nbytes > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, arena[:nbytes].contiguous(), root)]):
            req.wait()
    return None, sizes
