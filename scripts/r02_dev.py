"""Round-2 development timings: LBVH (compact nodes, self mode), EPA, pipeline stages.
    python scripts/r02_dev.py [bvh] [epa] [pipe]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from distance3d_b200 import _lib, aabb_tree, gjk, epa, random as R, pipeline

what = sys.argv[1:] or ["bvh", "epa", "pipe"]


def timed(label, fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-44s %9.3f ms (min %.3f)" % (label, float(np.mean(ts)), min(ts)), flush=True)
    return out, float(np.mean(ts))


if "bvh" in what:
    n = 1000000
    for name, scale in (("dense", 2.0), ("sparse", 2.0 * (n / 2000.0) ** (1.0 / 3.0))):
        cs = bench.make_capsules(n, scale)
        aabb = _lib.aabb_device(cs.device())
        bvh = aabb_tree.Lbvh(aabb)
        timed(name + " build", lambda: bvh.rebuild())
        pairs, count = bvh.overlap_self(ordered=False, packet=True, count_visits=True)
        v_other = bvh.visits()
        pu, cu = bvh.overlap_unique(count_visits=True)
        v_self = bvh.visits()
        print(name, "ordered-form pairs", count, "unique pairs", cu, "visits other", v_other, "self", v_self)
        assert 2 * cu + n == count
        buf = torch.empty((count, 2), dtype=torch.int32, device="cuda")
        timed(name + " other-mode all boxes, warp tiles", lambda: bvh.overlap_async(bvh.aabbs, buf, order=bvh.leaf_order(), packet=1))
        timed(name + " other-mode all boxes, per thread", lambda: bvh.overlap_async(bvh.aabbs, buf, order=bvh.leaf_order(), packet=0), reps=2)
        timed(name + " self-mode (unique pairs), warp tiles", lambda: bvh.overlap_unique_async(buf))
        timed(name + " ordered two-pass packet", lambda: bvh.overlap_self(packet=True), reps=2)
        del buf, pairs, pu, bvh
        torch.cuda.empty_cache()

if "epa" in what:
    rs = np.random.RandomState(85)
    n_pairs = 1000000
    cs = R.random_collider_set(rs, 2 * n_pairs, names=("mesh",), center_scale=0.7,
                               hull_vertices=(64, 256), hull_library=4096)
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(n_pairs, 2)
    dc = cs.device()
    g, t = timed("C3 gjk on %d hull pairs" % n_pairs, lambda: gjk.gjk_distance_batch(dc, pairs), reps=2)
    sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()
    pd = torch.from_numpy(pairs).cuda()[sel].contiguous(); Y = g.simplex[sel].contiguous()
    r, t = timed("C3 epa on %d pairs" % len(sel), lambda: epa.epa_batch(dc, pd, Y))
    print("C3 EPA pairs/s %.3e  iters %.1f  max_faces %.3f" % (len(sel) / t * 1e3, r.iters.double().mean().item(),
                                                              (r.status == 7).double().mean().item()))
    del cs, dc, g, r
    torch.cuda.empty_cache()

if "pipe" in what:
    n = 2000000
    rs = np.random.RandomState(84)
    scale = 0.33 * n ** (1.0 / 3.0)
    cs = R.random_collider_set(rs, n, names=R.PRIMITIVES + ("mesh",), center_scale=scale, hull_vertices=(10, 10))
    dc = cs.device()
    aabb, _ = timed("C5 aabb", lambda: _lib.aabb_device(dc))
    bvh, _ = timed("C5 bvh build (incl. alloc)", lambda: aabb_tree.Lbvh(aabb))
    (cand, count), _ = timed("C5 overlap_unique", lambda: bvh.overlap_unique(capacity=12 * n))
    g, _ = timed("C5 gjk distance (%d candidates)" % count, lambda: gjk.gjk_distance_batch(dc, cand))
    hits = torch.nonzero(g.dist == 0.0).flatten()
    pe, Y, npts = cand[hits].contiguous(), g.simplex[hits].contiguous(), g.n_points[hits].contiguous()
    r, t = timed("C5 epa (%d hits)" % len(hits), lambda: epa.epa_batch(dc, pe, Y, n_points=npts))
    print("C5 EPA pairs/s %.3e  iters %.1f  max_faces %.3f  bad simplex %.3f" % (
        len(hits) / t * 1e3, r.iters.double().mean().item(), (r.status == 7).double().mean().item(),
        (r.status == 8).double().mean().item()))
    res, t = timed("C5 pipeline.collide", lambda: pipeline.collide(dc, shard=False, candidate_capacity=12 * n), reps=3)
    print("C5 shapes/s %.3e" % (n / t * 1e3))
