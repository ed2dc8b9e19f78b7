import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distance3d_b200 import gjk, random as R, pack
from oracle import cpu_oracle as O
rs = np.random.RandomState(16)
cs = R.random_collider_set(rs, 4000, names=R.PRIMITIVES + ("mesh",))
pairs = R.random_pairs(rs, len(cs), 60000)
res = gjk.gjk_distance_batch(cs, pairs, dtype="f32").cpu()
ref = O.gjk_distance(cs, pairs, n_threads=8)
err = np.abs(res["dist"] - ref["dist"])
err[ref["status"] > 1] = 0
print("status f32", np.bincount(res["status"]), "ref", np.bincount(ref["status"]))
for thr in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1):
    print("err >", thr, (err > thr).sum())
bad = np.argsort(-err)[:12]
for k in bad:
    i, j = pairs[k]
    print(pack.TYPE_NAMES[cs.type[i]], pack.TYPE_NAMES[cs.type[j]], "err %.3e ref %.6f got %.6f iters ref %d got %d status %d npts %d/%d" % (
        err[k], ref["dist"][k], res["dist"][k], ref["iters"][k], res["iters"][k], res["status"][k], ref["n_points"][k], res["n_points"][k]))
tp = cs.type[pairs[:, 0]] * 10 + cs.type[pairs[:, 1]]
for t in np.unique(tp):
    m = tp == t
    print(pack.TYPE_NAMES[t // 10], pack.TYPE_NAMES[t % 10], "max err %.2e  frac>1e-4 %.4f" % (err[m].max(), (err[m] > 1e-4).mean()))
