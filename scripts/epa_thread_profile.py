"""Cycles per phase of the thread-per-pair EPA kernel (library built with -DEPA_PROFILE, warp clocks of lane 0):
    bash scripts/build_variant_epa.sh prof -DEPA_PROFILE
    D3D_B200_LIB=scripts/lib_epaprof.so python scripts/epa_thread_profile.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import _lib, aabb_tree, gjk, epa, random as R

NAMES = ["refill + init", "closest (tie scan)", "support + convergence", "vertex id", "visibility pass", "removal + edges", "new faces"]
n = int(os.environ.get("N", 500000))
rs = np.random.RandomState(84)
cs = R.random_collider_set(rs, n, names=R.PRIMITIVES + ("mesh",), center_scale=0.33 * n ** (1.0 / 3.0), hull_vertices=(10, 10))
dc = cs.device()
bvh = aabb_tree.Lbvh(_lib.aabb_device(dc))
cand, count = bvh.overlap_unique(capacity=12 * n)
g = gjk.gjk_distance_batch(dc, cand)
hits = torch.nonzero(g.dist == 0.0).flatten()
pe, Y, npts = cand[hits].contiguous(), g.simplex[hits].contiguous(), g.n_points[hits].contiguous()
os.environ["D3D_EPA_KERNEL"] = "thread"
epa.epa_batch(dc, pe, Y, n_points=npts)
out = (ctypes.c_uint64 * 16)()
_lib.lib().d3d_debug_epa_profile(out, 1)
r = epa.epa_batch(dc, pe, Y, n_points=npts)
_lib.lib().d3d_debug_epa_profile(out, 1)
v = np.array(list(out), dtype=np.float64)
tot = v[:7].sum()
print("C5 mix, %d pairs, %.1f iterations per pair; warp trips %d, lanes owning a pair per trip %.1f" % (
    len(pe), r.iters.double().mean().item(), int(v[9]), v[8] / max(v[9], 1)))
for k in range(7):
    print("   %-24s %5.1f %%   %8.0f cycles per warp trip" % (NAMES[k], 100 * v[k] / tot, v[k] / max(v[9], 1)))
