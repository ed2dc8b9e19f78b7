"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import gjk, epa, mpr, aabb_tree, pipeline, _lib, random as R, self_collision, broad_phase
from distance3d_b200.urdf import UrdfTransformManager

rs = np.random.RandomState(0)
cs = R.random_collider_set(rs, 600, names=R.PRIMITIVES + ("mesh", "cone"), center_scale=0.8, hull_vertices=(4, 100))
pairs = R.random_pairs(rs, len(cs), 3000)
g = gjk.gjk_distance_batch(cs, pairs)
gjk.gjk_intersection_batch(cs, pairs)
gjk.gjk_distance_batch(cs, pairs, dtype="f32")
sel = torch.nonzero((g.dist == 0) & (g.n_points == 4)).flatten()
epa.epa_batch(cs, torch.from_numpy(pairs).cuda()[sel], g.simplex[sel], want_faces=True)
mpr.mpr_batch(cs, pairs)
A = _lib.aabb_device(cs.device())
bvh = aabb_tree.Lbvh(A)
bvh.overlap_self(packet=True); bvh.overlap_self(packet=False)
for w in (0, 1, 2, 4, 16, 32):
    bvh.overlap_self(packet=w, ordered=False)
bvh.overlap(A[:100], capacity=8, ordered=False)
aabb_tree.brute_force_pairs(A[:200], A[200:500])
pipeline.collide(cs, shard=False)
data = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data")
tm = UrdfTransformManager()
tm.load_urdf(open(os.path.join(data, "robot_arm.urdf")).read(), mesh_path=data)
b = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
b.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
self_collision.RobotModel(tm, b).detect_batch(rs.uniform(-3, 3, size=(500, 6)))
_lib.debug_norm(rs.randn(5000, 3))
_lib.debug_vdiv(rs.randn(5000, 3), rs.randn(5000))
torch.cuda.synchronize()
print("sanitize run complete")
