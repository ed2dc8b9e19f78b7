"""Multi-GPU sharding: one process per GPU, `torch.distributed` for the plumbing.

The hot path has no exchange step: pairs, queries and joint configurations are
independent, so every rank works on a contiguous shard with a replicated collider
set / BVH.  The only collective is the variable-length
all-gather of result lists (contact lists, overlap pairs) when the caller wants them
in one place: an all-gather of the int64 counts, an exclusive scan, and one group of
exact-size point-to-point transfers into the slices of a single output buffer.
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


def all_gather_varlen(t, group=None, out=None):
    """Concatenate tensors whose first dimension differs between ranks; every rank gets all.

    Counts first (one small all-gather), then an exclusive scan gives every rank its slice of
    ONE output buffer and the slices travel as a single group of point-to-point transfers of
    exactly their size (NCCL: ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd over NVLink;
    gloo on CPU for the tests) - no padding to the largest rank, no staging copies.  Works on
    CPU and CUDA tensors.  Returns the concatenation in rank order and the per-rank counts;
    `out` (optional) is a buffer with room for the total that is re-used between calls.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t, [int(t.shape[0])]
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    t = t.contiguous()
    count = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts_t = torch.empty(ws, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(counts_t, count, group=group)
    counts = [int(c) for c in counts_t.cpu().tolist()]
    offsets = np.concatenate(([0], np.cumsum(counts)))
    total = int(offsets[-1])
    if out is None or out.shape[0] < total or out.dtype != t.dtype or tuple(out.shape[1:]) != tuple(t.shape[1:]):
        out = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    full = out[:total]
    ops = []
    for peer in range(ws):
        if peer == rank:
            continue
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, peer) if group else peer, group))
        if counts[peer]:
            ops.append(dist.P2POp(dist.irecv, full[offsets[peer]:offsets[peer + 1]],
                                  dist.get_global_rank(group, peer) if group else peer, group))
    full[offsets[rank]:offsets[rank + 1]].copy_(t)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full, counts


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


def overlap_unique_sharded(bvh, gather=True, out=None):
    """Self query with a replicated BVH: this rank walks its share of the leaves
    (`Lbvh.overlap_unique`: 128-leaf blocks dealt round-robin), so the ranks' lists of
    unordered pairs are disjoint; with gather=True every rank receives the whole list."""
    rank, ws = world()
    pairs, count = bvh.overlap_unique(rank, ws, out=out)
    if not gather:
        return pairs, count
    full, counts = all_gather_varlen(pairs)
    return full, int(sum(counts))


class PeerPairBuffer:
    """Symmetric pair buffer for the fused self query + all-gather (`d3d_bvh_overlap_self_gather`).

    Every rank allocates ``int32[world * segment_cap, 2]`` and ``uint64[world]`` through
    `torch.distributed._symmetric_memory` (CUDA virtual-memory handles exchanged between the
    processes of one NVLink / NVSwitch domain), so each GPU holds device pointers to the buffers
    of all GPUs.  The traversal kernel of rank r stores its pairs into segment r of EVERY buffer;
    after `finish()` (a barrier on the stream) each rank holds the whole list as `world`
    segments: ``pairs[r * segment_cap : r * segment_cap + counts[r]]``.
    """

    def __init__(self, segment_cap, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.segment_cap = int(segment_cap)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.pairs = symm.empty((self.world * self.segment_cap, 2), dtype=torch.int32, device=dev)
        self.counts = symm.empty(self.world, dtype=torch.int64, device=dev)
        self.counts.zero_()
        self._h_pairs = symm.rendezvous(self.pairs, self.group)
        self._h_counts = symm.rendezvous(self.counts, self.group)
        self.local_count = torch.zeros(1, dtype=torch.int64, device=dev)

    @property
    def peer_pairs_ptr(self):
        return self._h_pairs.buffer_ptrs_dev

    @property
    def peer_counts_ptr(self):
        return self._h_counts.buffer_ptrs_dev

    def finish(self):
        """All ranks' stores have landed when this returns control to the stream."""
        self._h_pairs.barrier()

    def segments(self):
        """List of the ranks' pair lists (views; synchronises the host to read the counts)."""
        counts = [int(c) for c in self.counts.cpu().tolist()]
        return [self.pairs[r * self.segment_cap: r * self.segment_cap + min(c, self.segment_cap)]
                for r, c in enumerate(counts)], counts


def overlap_unique_fused_gather(bvh, buf, count_visits=False):
    """Self query of a replicated BVH with the all-gather fused into the traversal kernel:
    this rank walks its share of the leaves and stores its pairs into every GPU's
    :class:`PeerPairBuffer` over NVLink.  Returns after the cross-GPU barrier was enqueued."""
    import ctypes
    from . import _lib
    from ._lib import c_i64, ptr
    buf.finish()   # nobody is still reading the previous round's segments
    _lib._check(_lib.lib().d3d_bvh_overlap_self_gather(
        ptr(bvh.workspace), c_i64(bvh.n), ctypes.c_int(buf.rank), ctypes.c_int(buf.world),
        ctypes.c_void_p(buf.peer_pairs_ptr), c_i64(buf.segment_cap), ctypes.c_void_p(buf.peer_counts_ptr),
        ptr(buf.local_count), ptr(bvh._visits) if count_visits else None, _lib.stream_ptr()))
    buf.finish()
    return buf
