"""Single-pass LBVH query time vs packet width (C2 dense / constant density)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distance3d_b200 import _lib, aabb_tree
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
for name, scale in (("dense", 2.0), ("constant_density", 2.0 * (n / 2000.0) ** (1.0 / 3.0))):
    aabb = _lib.aabb_device(bench.make_capsules(n, scale).device())
    bvh = aabb_tree.Lbvh(aabb)
    ref, count = bvh.overlap_self()
    buf = torch.empty((count, 2), dtype=torch.int32, device="cuda")
    for width in (0, 2, 4, 8, 16, 32):
        pairs, c = bvh.overlap_self(packet=width, ordered=False, out=buf, count_visits=True)
        assert c == count
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); bvh.overlap_async(bvh.aabbs, buf, order=bvh.leaf_order(), packet=width); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print("%-17s width %2d  %8.3f ms  %.3e pairs/s  node fetches %d" % (name, max(width, 1), min(ts), count / min(ts) * 1e3, bvh.visits()))
    del bvh, buf, ref
