"""C1 headline timing only: python scripts/gjk_c1.py [pairs]  (D3D_B200_LIB selects a variant build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from distance3d_b200 import gjk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
cs, pairs = bench.make_workload(84, n)
dc = cs.device(); pd = torch.from_numpy(pairs).cuda()
out = gjk.gjk_distance_batch(dc, pd)
for _ in range(3): gjk.gjk_distance_batch(dc, pd, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): gjk.gjk_distance_batch(dc, pd, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
chk = float(out.dist[out.status <= 1].sum().item())
print("%-28s C1 %d pairs: %.3f ms  %.4e pairs/s  iters %.3f  checksum %.10e" % (
    os.path.basename(os.environ.get("D3D_B200_LIB", "default")), n, ms, n / ms * 1e3,
    out.iters.double().mean().item(), chk), flush=True)
hit, _, _ = gjk.gjk_intersection_batch(dc, pd)
for _ in range(2): gjk.gjk_intersection_batch(dc, pd)
torch.cuda.synchronize(); e0.record()
for _ in range(5): gjk.gjk_intersection_batch(dc, pd)
e1.record(); torch.cuda.synchronize()
print("   intersection: %.4e pairs/s" % (n / (e0.elapsed_time(e1) / 5) * 1e3))
