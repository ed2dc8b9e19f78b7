import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from distance3d_b200 import _lib, aabb_tree
n = 1000000
aabb = _lib.aabb_device(bench.make_capsules(n, 2.0 * (n / 2000.0) ** (1.0 / 3.0)).device())
bvh = aabb_tree.Lbvh(aabb)
for _ in range(3): bvh.rebuild()
torch.cuda.synchronize()
