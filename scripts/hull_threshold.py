"""GJK on convex hulls: thread-per-pair vs warp-per-pair as a function of hull size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import gjk, random as R
lo, hi = int(sys.argv[1]), int(sys.argv[2])
n = 200000
rs = np.random.RandomState(3)
cs = R.random_collider_set(rs, 2 * n, names=("mesh",), center_scale=0.7, hull_vertices=(lo, hi), hull_library=4096)
pairs = np.arange(2 * n, dtype=np.int32).reshape(n, 2)
dc = cs.device(); pd = torch.from_numpy(pairs).cuda()
out = gjk.gjk_distance_batch(dc, pd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): gjk.gjk_distance_batch(dc, pd, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("hull_max=%s verts %d-%d: %.1f Mpairs/s iters %.1f" % (os.environ.get("D3D_THREAD_HULL_MAX", "16"), lo, hi, n / ms / 1e3, out.iters.double().mean().item()))
