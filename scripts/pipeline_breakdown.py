"""Stage-by-stage timing of the C5 pipeline (LBVH + GJK + EPA) on one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import _lib, aabb_tree, gjk, epa, random as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
rs = np.random.RandomState(84)
scale = 0.33 * n ** (1.0 / 3.0)
cs = R.random_collider_set(rs, n, names=R.PRIMITIVES + ("mesh",), center_scale=scale, hull_vertices=(8, 32))
dc = cs.device()

def timed(label, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print("%-28s %9.2f ms" % (label, e0.elapsed_time(e1) / reps))
    return out

aabb = timed("aabb", lambda: _lib.aabb_device(dc))
bvh = timed("bvh build (incl. alloc)", lambda: aabb_tree.Lbvh(aabb))
timed("bvh rebuild", lambda: bvh.rebuild())
pairs, count = timed("overlap_self", lambda: bvh.overlap_self())
cand = timed("filter i<j", lambda: pairs[pairs[:, 0] < pairs[:, 1]].contiguous())
print("candidates", cand.shape[0], "overlaps", count)
g = timed("gjk distance", lambda: gjk.gjk_distance_batch(dc, cand))
hits = timed("select hits", lambda: torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten())
print("epa pairs", hits.numel(), "mean gjk iters", g.iters.double().mean().item())
pe, Y = cand[hits].contiguous(), g.simplex[hits].contiguous()
r = timed("epa", lambda: epa.epa_batch(dc, pe, Y))
print("epa iters mean", r.iters.double().mean().item(), "max_faces rate", (r.status == 7).double().mean().item())
types = dc.type[pe.long()]
key = (types[:, 0] * 10 + types[:, 1]).cpu().numpy()
