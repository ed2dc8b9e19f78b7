"""Development: thread-per-pair EPA vs warp EPA on the C3 / C5 workloads (deferred fraction, ms)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import _lib, aabb_tree, gjk, epa, random as R


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, min(ts)


def run(label, dc, pe, Y, npts):
    os.environ.pop("D3D_EPA_KERNEL", None)
    r, t = timed(lambda: epa.epa_batch(dc, pe, Y, n_points=npts))
    os.environ["D3D_EPA_KERNEL"] = "warp"
    w, tw = timed(lambda: epa.epa_batch(dc, pe, Y, n_points=npts))
    os.environ.pop("D3D_EPA_KERNEL", None)
    ok = (w.status != 7)
    same = bool((r.status == w.status).all()) and bool((r.mtv[ok] == w.mtv[ok]).all())
    print("%-28s %8d pairs  thread+warp %8.3f ms  warp only %8.3f ms  deferred %.4f  iters %.1f  identical %s"
          % (label, len(pe), t, tw, int(r.deferred[0]) / max(len(pe), 1), w.iters.double().mean().item(), same), flush=True)


what = sys.argv[1:] or ["c3", "c5", "types"]
if "c3" in what:
    rs = np.random.RandomState(85)
    n_pairs = 400000
    cs = R.random_collider_set(rs, 2 * n_pairs, names=("mesh",), center_scale=0.7, hull_vertices=(64, 256), hull_library=4096)
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(n_pairs, 2)
    dc = cs.device()
    g = gjk.gjk_distance_batch(dc, pairs)
    sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()
    run("C3 hulls 64-256", dc, torch.from_numpy(pairs).cuda()[sel].contiguous(), g.simplex[sel].contiguous(), None)
    del cs, dc, g
if "c5" in what or "types" in what:
    n = int(os.environ.get("N", 500000))
    rs = np.random.RandomState(84)
    scale = 0.33 * n ** (1.0 / 3.0)
    cs = R.random_collider_set(rs, n, names=R.PRIMITIVES + ("mesh",), center_scale=scale, hull_vertices=(10, 10))
    dc = cs.device()
    bvh = aabb_tree.Lbvh(_lib.aabb_device(dc))
    cand, count = bvh.overlap_unique(capacity=12 * n)
    g = gjk.gjk_distance_batch(dc, cand)
    hits = torch.nonzero(g.dist == 0.0).flatten()
    pe, Y, npts = cand[hits].contiguous(), g.simplex[hits].contiguous(), g.n_points[hits].contiguous()
    if "c5" in what:
        run("C5 mix", dc, pe, Y, npts)
    if "types" in what:
        types = torch.from_numpy(cs.type.astype(np.int64)).cuda()
        for ta in range(6):
            m = (types[pe[:, 0].long()] == ta) & (types[pe[:, 1].long()] == ta) & (npts == 4)
            idx = torch.nonzero(m).flatten()
            if len(idx) >= 100:
                run("C5 type %d x %d" % (ta, ta), dc, pe[idx].contiguous(), Y[idx].contiguous(), npts[idx].contiguous())
