"""Multi-GPU sharding: one process per GPU, `torch.distributed` for the plumbing.

The hot path has no exchange step: pairs, queries and joint configurations are
independent, so every rank works on a contiguous shard with a replicated collider
set / BVH.  The only collective is the variable-length
all-gather of result lists (contact lists, overlap pairs) when the caller wants them
in one place: an all-gather of the int64 counts followed by a padded all-gather.
"""
import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous, balanced shard [begin, end) of n work items for `rank`."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_counts(n, world_size):
    return [shard_range(n, r, world_size)[1] - shard_range(n, r, world_size)[0]
            for r in range(world_size)]


def all_gather_varlen(t, group=None):
    """Concatenate tensors whose first dimension differs between ranks.

    Works on CPU tensors (gloo) and CUDA tensors (NCCL).  Returns the
    concatenation in rank order and the list of per-rank counts.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t, [int(t.shape[0])]
    ws = dist.get_world_size(group)
    count = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(count) for _ in range(ws)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[:t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0), counts


def gjk_distance_sharded(colliders, pairs, gather=False, **kwargs):
    """GJK distance over this rank's shard of `pairs` (int32[P,2], identical on all
    ranks).  Returns (begin, end, GjkResult); with gather=True the distances of all
    ranks are all-gathered (every rank gets the full float64[P])."""
    from . import gjk
    begin, end = shard_range(len(pairs))
    res = gjk.gjk_distance_batch(colliders, pairs[begin:end], **kwargs)
    if gather:
        full, _ = all_gather_varlen(res.dist)
        return begin, end, res, full
    return begin, end, res


def overlap_sharded(bvh, query, gather=True):
    """All-overlap query with a replicated BVH: this rank traverses its shard of the
    query boxes; pair lists (tree index, query index) are all-gathered."""
    import torch
    begin, end = shard_range(int(query.shape[0]))
    pairs, count = bvh.overlap(query[begin:end])
    if count:
        pairs = pairs.clone()
        pairs[:, 1] += begin
    if not gather:
        return pairs, count
    full, counts = all_gather_varlen(pairs)
    return full, int(sum(counts))
