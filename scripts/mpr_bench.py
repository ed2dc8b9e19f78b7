import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import mpr, random as R
n = 1 << 20
rs = np.random.RandomState(2)
cs = R.random_collider_set(rs, 2 * n, names=R.PRIMITIVES, center_scale=0.7)
pairs = torch.from_numpy(np.arange(2 * n, dtype=np.int32).reshape(n, 2)).cuda()
dc = cs.device()
for pen in (True, False):
    mpr.mpr_batch(dc, pairs, penetration=pen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = mpr.mpr_batch(dc, pairs, penetration=pen)
    e1.record(); torch.cuda.synchronize()
    print("mpr penetration=%s: %.1f Mpairs/s hit %.2f" % (pen, n / (e0.elapsed_time(e1) / 3) / 1e3, out["hit"].double().mean().item()))
