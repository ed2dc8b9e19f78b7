"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distance3d_b200 import gjk, epa, mpr, aabb_tree, pipeline, _lib, random as R, self_collision, broad_phase
from distance3d_b200.urdf import UrdfTransformManager

N_COL, N_PAIRS = int(os.environ.get("D3D_SAN_COLLIDERS", 600)), int(os.environ.get("D3D_SAN_PAIRS", 3000))
rs = np.random.RandomState(0)
cs = R.random_collider_set(rs, N_COL, names=R.PRIMITIVES + ("mesh", "cone"), center_scale=0.8, hull_vertices=(4, 100))
pairs = R.random_pairs(rs, len(cs), N_PAIRS)
g = gjk.gjk_distance_batch(cs, pairs)
gjk.gjk_intersection_batch(cs, pairs)
gjk.gjk_distance_batch(cs, pairs, dtype="f32")
sel = torch.nonzero(g.dist == 0).flatten()
epa.epa_batch(cs, torch.from_numpy(pairs).cuda()[sel], g.simplex[sel], want_faces=True, n_points=g.n_points[sel])
epa.epa_batch(cs, torch.from_numpy(pairs).cuda()[sel], g.simplex[sel], max_faces=48, n_points=g.n_points[sel])
# the thread-per-pair EPA kernel (batches this small go to the warp kernel unless forced)
os.environ["D3D_EPA_KERNEL"] = "thread"
epa.epa_batch(cs, torch.from_numpy(pairs).cuda()[sel], g.simplex[sel], n_points=g.n_points[sel])
Yd = g.simplex[sel].clone()
Yd[::5, 1] = Yd[::5, 0]; Yd[2::7, 2] = Yd[2::7, 3] + 1e-9   # shared vertex ids, near pairs
epa.epa_batch(cs, torch.from_numpy(pairs).cuda()[sel], Yd)
os.environ.pop("D3D_EPA_KERNEL")
gjk.gjk_intersection_libccd_batch(cs, pairs)
# MeshGraph colliders: hill climbing in the thread kernel, the warp kernel, EPA and MPR
mg = R.random_meshgraph_set(rs, 6, 60, 60, hull_vertices=(8, 150), center_scale=0.8)
mp = R.random_pairs(rs, len(mg), 600)
gm = gjk.gjk_distance_batch(mg, mp)
gjk.gjk_intersection_batch(mg, mp)
selm = torch.nonzero(gm.dist == 0).flatten()
epa.epa_batch(mg, torch.from_numpy(mp).cuda()[selm], gm.simplex[selm], n_points=gm.n_points[selm])
os.environ["D3D_EPA_KERNEL"] = "thread"   # MeshGraph pairs are handed to the warp kernel
epa.epa_batch(mg, torch.from_numpy(mp).cuda()[selm], gm.simplex[selm], n_points=gm.n_points[selm])
os.environ.pop("D3D_EPA_KERNEL")
mpr.mpr_batch(mg, mp)
mpr.mpr_batch(cs, pairs)
A = _lib.aabb_device(cs.device())
bvh = aabb_tree.Lbvh(A)
bvh.overlap_self(packet=True); bvh.overlap_self(packet=False)
for w in (0, 1, 2, 4, 16, 32):
    bvh.overlap_self(packet=w, ordered=False)
bvh.overlap(A[:100], capacity=8, ordered=False)
for w in (0, 1, 2, 32):
    bvh.overlap_unique(packet=w)
bvh.overlap_unique(1, 3)
aabb_tree.brute_force_pairs(A[:200], A[200:500])
pipeline.collide(cs, shard=False)
data = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "data")
tm = UrdfTransformManager()
tm.load_urdf(open(os.path.join(data, "robot_arm.urdf")).read(), mesh_path=data)
b = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
b.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
self_collision.RobotModel(tm, b).detect_batch(rs.uniform(-3, 3, size=(500, 6)))
# branched robot: ordered replay of the reference's detect loop
tm2 = UrdfTransformManager()
tm2.load_urdf(open(os.path.join(data, "robot_branched.urdf")).read(), mesh_path=data)
b2 = broad_phase.BoundingVolumeHierarchy(tm2, "robot_branched")
b2.fill_tree_with_colliders(tm2, fill_self_collision_whitelists=True)
self_collision.RobotModel(tm2, b2).detect_batch(rs.uniform(-3, 3, size=(300, 5)))
# wire records and the host-buffer stream
from distance3d_b200 import stream as d3stream, hydroelastic_contact as hc
pipe = d3stream.GjkDistanceStream(len(cs), len(pairs), cs.n_vertices, slots=2)
pipe.result(pipe.submit(d3stream.pin_batch(cs, pairs, wire=True)))
# tetrahedron pairs
tp = rs.randn(200, 4, 3) * 0.3 + rs.randn(200, 1, 3)
hc.find_contact_pairs(tp[:100], rs.rand(100, 4), tp[100:], rs.rand(100, 4))
_lib.debug_norm(rs.randn(5000, 3))
_lib.debug_vdiv(rs.randn(5000, 3), rs.randn(5000))
torch.cuda.synchronize()
print("sanitize run complete")
