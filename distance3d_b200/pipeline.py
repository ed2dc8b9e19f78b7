"""Full collision pipeline on the device: broad phase -> GJK -> EPA.

Batched counterpart of the loop in the reference's examples
(examples/visualizations/vis_capsules_benchmark.py:38-50): build the BVH over all
collider AABBs, take every pair with overlapping boxes, run the narrow phase on the
candidates and, for the intersecting ones, the penetration query.  With several ranks
(one process per GPU) every rank holds a replica of the BVH and handles a contiguous
shard of the query boxes; nothing is exchanged until the
caller asks for the gathered contact list.
"""
import numpy as np

from . import _lib, aabb_tree, epa as _epa, gjk as _gjk, parallel


class PipelineResult:
    """Candidates and contacts of one :func:`collide` call (device tensors)."""

    def __init__(self, n_overlaps, candidates, gjk, hits, epa):
        self.n_overlaps = n_overlaps    # ordered AABB overlaps found by the traversal (incl. i == j)
        self.candidates = candidates    # int32[C,2], i < j
        self.gjk = gjk                  # GjkResult over the candidates
        self.hits = hits                # int64[H] indices into candidates with distance 0
        self.epa = epa                  # EpaResult over the hits with a 4-point simplex (or None)
        self.epa_index = None           # int64[E] indices into candidates that went through EPA


def collide(colliders, penetration=True, distance_threshold=None, shard=True):
    """Broad phase + narrow phase for all colliders of a packed set.

    Returns a :class:`PipelineResult`.  `distance_threshold` is only used to report
    `near` pairs by the caller; all candidates get an exact GJK distance.
    """
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    aabb = _lib.aabb_device(dc)
    bvh = aabb_tree.Lbvh(aabb)
    order = bvh.leaf_order()
    if shard:
        begin, end = parallel.shard_range(dc.n)
        order = order[begin:end].contiguous()
    pairs, count = bvh.overlap(aabb, order=order)
    # every unordered pair once (the traversal reports both orientations and (i, i))
    # (with several ranks the one that owns the larger index as query keeps the pair)
    keep = pairs[:, 0] < pairs[:, 1]
    candidates = pairs[keep].contiguous()
    g = _gjk.gjk_distance_batch(dc, candidates)
    hits = torch.nonzero(g.dist == 0.0).flatten()
    res = PipelineResult(count, candidates, g, hits, None)
    if penetration and hits.numel():
        full = hits[g.n_points[hits] == 4]
        res.epa_index = full
        res.epa = _epa.epa_batch(dc, candidates[full], g.simplex[full])
    return res
