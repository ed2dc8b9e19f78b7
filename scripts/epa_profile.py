"""Cycles per phase of the EPA iteration (library built with -DEPA_PROFILE):
    bash scripts/build_variant_epa.sh prof -DEPA_PROFILE
    D3D_B200_LIB=scripts/lib_epaprof.so python scripts/epa_profile.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import _lib, aabb_tree, gjk, epa, random as R

NAMES = ["setup", "A closest face", "B support", "C + visibility", "D edge loop", "D apply perm", "E new faces", "epilogue"]


def report(label, n_pairs):
    out = (ctypes.c_uint64 * 16)()
    _lib.lib().d3d_debug_epa_profile(out, 1)
    v = np.array(list(out), dtype=np.float64)
    tot = v[:8].sum()
    print("== %s: %d pairs, %.0f cycles per pair (sum over the warp's phases)" % (label, n_pairs, tot / n_pairs))
    for k in range(8):
        print("   %-16s %5.1f %%  %8.0f cycles/pair" % (NAMES[k], 100 * v[k] / tot, v[k] / n_pairs))
    it = max(v[8], 1)
    print("   expanding iterations/pair %.2f  faces/iter %.1f  visible/iter %.2f  removed/iter %.2f  loose/iter %.2f"
          % (v[8] / n_pairs, v[9] / it, v[10] / it, v[11] / it, v[12] / it))
    print("   cycles per expanding iteration: edge loop %.0f (%.0f per removed face), new faces %.0f, support %.0f"
          % (v[4] / it, v[4] / max(v[11], 1), v[6] / it, v[2] / it), flush=True)


rs = np.random.RandomState(85)
n_pairs = 400000
cs = R.random_collider_set(rs, 2 * n_pairs, names=("mesh",), center_scale=0.7, hull_vertices=(64, 256), hull_library=4096)
pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(n_pairs, 2)
dc = cs.device()
g = gjk.gjk_distance_batch(dc, pairs)
sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()
pd = torch.from_numpy(pairs).cuda()[sel].contiguous(); Y = g.simplex[sel].contiguous()
_lib.lib().d3d_debug_epa_profile((ctypes.c_uint64 * 16)(), 1)
r = epa.epa_batch(dc, pd, Y)
report("C3 hulls", len(sel))
del cs, dc, g, r

n = 500000
rs = np.random.RandomState(84)
scale = 0.33 * n ** (1.0 / 3.0)
cs = R.random_collider_set(rs, n, names=R.PRIMITIVES + ("mesh",), center_scale=scale, hull_vertices=(10, 10))
dc = cs.device()
bvh = aabb_tree.Lbvh(_lib.aabb_device(dc))
cand, count = bvh.overlap_unique(capacity=12 * n)
g = gjk.gjk_distance_batch(dc, cand)
hits = torch.nonzero(g.dist == 0.0).flatten()
pe, Y, npts = cand[hits].contiguous(), g.simplex[hits].contiguous(), g.n_points[hits].contiguous()
_lib.lib().d3d_debug_epa_profile((ctypes.c_uint64 * 16)(), 1)
r = epa.epa_batch(dc, pe, Y, n_points=npts)
report("C5 mix", int((npts == 4).sum()))
# per type pair
types = torch.from_numpy(cs.type.astype(np.int64)).cuda()
for ta in range(len(R.PRIMITIVES) + 1):
    m = (types[pe[:, 0].long()] == ta) & (types[pe[:, 1].long()] == ta) & (npts == 4)
    idx = torch.nonzero(m).flatten()
    if len(idx) < 100:
        continue
    _lib.lib().d3d_debug_epa_profile((ctypes.c_uint64 * 16)(), 1)
    epa.epa_batch(dc, pe[idx].contiguous(), Y[idx].contiguous(), n_points=npts[idx].contiguous())
    report("C5 type %d x %d" % (ta, ta), len(idx))
