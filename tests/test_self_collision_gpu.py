"""GPU: URDF forward kinematics + batched self-collision (BASELINE config 4) and the
scalar BoundingVolumeHierarchy / self_collision API, against the reference's pinned
numbers and its own outputs on random joint configurations (tests/golden)."""
import os

import numpy as np
import pytest

from distance3d_b200 import broad_phase, colliders, self_collision
from distance3d_b200.urdf import UrdfTransformManager, TransformManager
from util import GOLDEN

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_robot():
    tm = UrdfTransformManager()
    with open(os.path.join(DATA, "robot_arm.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=DATA)
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
    bvh.fill_tree_with_colliders(tm, make_artists=False, fill_self_collision_whitelists=True)
    return tm, bvh


def test_bvh_reference_pins():
    # distance3d/test/test_broad_phase.py:11-29
    tm, bvh = load_robot()
    assert len(bvh.get_artists()) == 0
    assert len(bvh.get_collider_frames()) == 8 and len(bvh.get_colliders()) == 8
    tm.set_joint("joint1", 3.1415926535)
    bvh.update_collider_poses()
    box = colliders.Box(np.eye(4), np.array([1, 1, 1], dtype=float))
    assert len(bvh.aabb_overlapping_colliders(box)) == 5


def test_self_collision_reference_pins():
    # distance3d/test/test_self_collision.py:9-42
    tm, bvh = load_robot()
    assert np.count_nonzero(list(self_collision.detect(bvh).values())) == 0
    assert not self_collision.detect_any(bvh)
    for j, v in (("joint2", 1.57), ("joint3", 1.57), ("joint5", 1.93)):
        tm.set_joint(j, v)
    bvh.update_collider_poses()
    assert np.count_nonzero(list(self_collision.detect(bvh).values())) == 0
    tm.set_joint("joint5", 2.05)
    bvh.update_collider_poses()
    assert np.count_nonzero(list(self_collision.detect(bvh).values())) == 3
    assert self_collision.detect_any(bvh)


def test_batched_fk_and_contact_masks_vs_reference_outputs():
    g = np.load(os.path.join(GOLDEN, "self_collision.npz"))
    tm, bvh = load_robot()
    model = self_collision.RobotModel(tm, bvh)
    assert [str(f) for f in g["frames"]] == model.frames
    assert len(model.pattern) == 17          # 28 - 11 white-listed pairs
    poses = model.forward_kinematics(g["q"]).cpu().numpy()
    assert np.max(np.abs(poses - g["poses"])) < 1e-12
    mask, n_cand = model.detect_batch(g["q"])
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    assert mask[1].sum() == 0 and mask[2].sum() == 3
    # chunked evaluation gives the same answer
    mask2, _ = model.detect_batch(np.tile(g["q"], (3, 1)), chunk=64)
    assert np.array_equal(mask2.cpu().numpy(), np.tile(g["mask"], (3, 1)))


@pytest.mark.parametrize("urdf,name", [("robot_arm.urdf", "robot_arm"), ("robot_branched.urdf", "robot_branched")])
def test_fk_with_shared_chain_prefixes_equals_per_frame_chains(urdf, name):
    """d3d_fk_urdf_tree (one thread per configuration, common chain prefixes once) performs, for every
    frame, the operations of d3d_fk_urdf in the same order: poses identical bit for bit."""
    tm = UrdfTransformManager()
    tm.load_urdf(open(os.path.join(DATA, urdf)).read(), mesh_path=DATA)
    b = broad_phase.BoundingVolumeHierarchy(tm, name)
    b.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
    model = self_collision.RobotModel(tm, b)
    q = np.random.RandomState(3).uniform(-3.2, 3.2, size=(5000, model.n_joints))
    tree = model.forward_kinematics(q).cpu().numpy()
    flat = model.forward_kinematics(q, shared_prefixes=False).cpu().numpy()
    assert np.array_equal(tree, flat)
    assert np.array_equal(tree[:, :, 3], np.broadcast_to([0.0, 0.0, 0.0, 1.0], tree[:, :, 3].shape))


def test_branched_robot_asymmetric_whitelists_equal_reference_detect():
    """Torso with two arms and a head: the reference's white-lists are asymmetric there and its
    detect() result depends on the candidate order of its AABB tree; the device replays that
    order (d3d_detect_ordered).  Fixture: reference detect() over 300 joint configurations."""
    g = np.load(os.path.join(GOLDEN, "self_collision_branched.npz"))
    tm = UrdfTransformManager()
    with open(os.path.join(DATA, "robot_branched.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=DATA)
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "robot_branched")
    bvh.fill_tree_with_colliders(tm, make_artists=False, fill_self_collision_whitelists=True)
    model = self_collision.RobotModel(tm, bvh)
    assert model.frames == [str(f) for f in g["frames"]] and not model.symmetric
    assert list(model.joint_names) == [str(j) for j in g["joints"]]
    mask, _ = model.detect_batch(g["q"])
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    mask2, _ = model.detect_batch(np.tile(g["q"], (2, 1)), chunk=128)
    assert np.array_equal(mask2.cpu().numpy(), np.tile(g["mask"], (2, 1)))
    for b in (0, 5, 17, 123):  # scalar API on the same configurations
        for name, v in zip(model.joint_names, g["q"][b]):
            tm.set_joint(name, v)
        bvh.update_collider_poses()
        contacts = self_collision.detect(bvh)
        assert [int(contacts[f]) for f in model.frames] == list(g["mask"][b])
        assert self_collision.detect_any(bvh) == bool(g["any"][b])


def test_bvh_from_colliders_translation_invariance():
    # distance3d/test/test_broad_phase.py:32-101
    from distance3d_b200 import random as d3random
    from distance3d_b200._transforms import transform_from
    tm = TransformManager()
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "origin")
    rs = np.random.RandomState(232)
    poses = {}
    box2origin, size = d3random.rand_box(rs, center_scale=1.0)
    bvh.add_collider("box", colliders.Box(box2origin, size)); poses["box"] = box2origin
    center, radius = d3random.rand_sphere(rs, center_scale=1.0)
    bvh.add_collider("sphere", colliders.Sphere(center, radius))
    poses["sphere"] = transform_from(np.eye(3), center)
    c2o, radius, height = d3random.rand_capsule(rs, center_scale=1.0)
    bvh.add_collider("capsule", colliders.Capsule(c2o, radius, height)); poses["capsule"] = c2o
    c2o, radius, length = d3random.rand_cylinder(rs, center_scale=1.0)
    bvh.add_collider("cylinder", colliders.Cylinder(c2o, radius, length)); poses["cylinder"] = c2o
    e2o, radii = d3random.rand_ellipsoid(rs, center_scale=1.0)
    bvh.add_collider("ellipsoid", colliders.Ellipsoid(e2o, radii)); poses["ellipsoid"] = e2o
    for frame, pose in poses.items():
        tm.add_transform(frame, "origin", pose)

    def overlaps():
        return {f: sorted(bvh.aabb_overlapping_colliders(c, whitelist=(f,)).keys())
                for f, c in bvh.colliders_.items()}

    before = overlaps()
    for frame, pose in poses.items():
        moved = np.copy(pose)
        moved[2, 3] -= 1.0
        tm.add_transform(frame, "origin", moved)
    bvh.update_collider_poses()
    assert overlaps() == before
    pairs = bvh.aabb_overlapping_with_other_bvh(bvh)
    pairs1 = [((f, c), (f2, c2)) for (f, c), (f2, c2) in pairs if f != f2]
    pairs2 = bvh.aabb_overlapping_with_self()
    assert sorted((a[0], b[0]) for a, b in pairs1) == sorted((a[0], b[0]) for a, b in pairs2)


MESH_URDF = """<?xml version="1.0"?>
<robot name="mesh_bot">
  <link name="base"><collision><geometry><mesh filename="part.stl" scale="1 2 0.5"/></geometry></collision></link>
  <link name="arm"><collision><geometry><mesh filename="part.stl"/></geometry></collision></link>
  <link name="tip"><collision><origin xyz="0.2 0 0" rpy="0 0.3 0"/>
    <geometry><mesh filename="part.stl" scale="0.5 0.5 0.5"/></geometry></collision></link>
  <link name="ghost"><collision><geometry><mesh filename="missing.stl"/></geometry></collision></link>
  <joint name="j1" type="revolute"><parent link="base"/><child link="arm"/>
    <origin xyz="0 0 2.5" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
  <joint name="j2" type="revolute"><parent link="arm"/><child link="tip"/>
    <origin xyz="0 0 2.5" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
  <joint name="j3" type="fixed"><parent link="tip"/><child link="ghost"/></joint>
</robot>"""


def test_urdf_mesh_colliders_through_the_bvh(tmp_path):
    """<mesh> geometry -> io.load_mesh -> MeshGraph -> BVH / self-collision; the
    missing file is a warning (broad_phase.py:76-81), results agree with the CPU oracle."""
    from distance3d_b200 import io, mesh, pack
    from oracle import cpu_oracle as O
    rs = np.random.RandomState(9)
    V = rs.randn(40, 3)
    V -= V.mean(axis=0)
    io.save_stl(str(tmp_path / "part.stl"), V, mesh.make_convex_mesh(V), binary=True)
    tm = UrdfTransformManager()
    tm.load_urdf(MESH_URDF, mesh_path=str(tmp_path))
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "mesh_bot")
    with pytest.warns(UserWarning, match="missing.stl"):
        bvh.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
    frames = ["collision:base/0", "collision:arm/0", "collision:tip/0"]
    assert sorted(bvh.get_collider_frames()) == sorted(frames)
    assert all(isinstance(c, colliders.MeshGraph) for c in bvh.get_colliders())
    seen = set()
    for q1, q2 in ((0.0, 0.0), (2.6, 2.9), (1.2, -2.0), (3.0, 3.0), (-2.7, -2.8)):
        tm.set_joint("j1", q1)
        tm.set_joint("j2", q2)
        bvh.update_collider_poses()
        got = self_collision.detect(bvh)
        # only base <-> tip is not white-listed (parent / child links are skipped)
        cs = pack.pack_colliders([bvh.colliders_["collision:base/0"], bvh.colliders_["collision:tip/0"]])
        hit = bool(O.gjk_intersection(cs, np.array([[0, 1]], dtype=np.int32))["hit"][0])
        assert got["collision:base/0"] == hit and got["collision:tip/0"] == hit
        assert got["collision:arm/0"] is False or got["collision:arm/0"] == False  # noqa: E712
        assert self_collision.detect_any(bvh) == hit
        seen.add(hit)
    assert seen == {True, False}


def test_mesh_robot_reference_pins_0_2_4():
    """distance3d/test/test_self_collision.py:44-76: sphere + three cone meshes (URDF <mesh>,
    binary STL) at joint angles 0 / 1.5 / 2.7 -> 0 / 2 / 4 colliders in contact.  The fixtures
    are re-authored (tests/data/simple_mechanism.urdf, cone.stl); the real reference run on
    them gives the same 0 / 2 / 4 and the same frames (checked when the fixtures were made)."""
    tm = UrdfTransformManager()
    with open(os.path.join(DATA, "simple_mechanism.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=DATA)
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "simple_mechanism")
    bvh.fill_tree_with_colliders(tm, make_artists=False, fill_self_collision_whitelists=True)
    frames = ["collision:cone%d/cone1" % k for k in (1, 2, 3, 4)]
    assert list(bvh.colliders_.keys()) == frames
    cone = bvh.colliders_[frames[1]]
    assert isinstance(cone, colliders.MeshGraph) and cone.vertices.shape == (64, 3) and len(cone.triangles) == 124
    expect = {0.0: [0, 0, 0, 0], 1.5: [1, 0, 0, 1], 2.7: [1, 1, 1, 1]}
    for q, mask in expect.items():
        for j in ("joint1", "joint2", "joint3"):
            tm.set_joint(j, q)
        bvh.update_collider_poses()
        contacts = self_collision.detect(bvh)
        assert [int(contacts[f]) for f in frames] == mask
        assert self_collision.detect_any(bvh) == bool(sum(mask))
    # the batched model handles MeshGraph colliders too (shared vertices / adjacency, one pose per configuration)
    model = self_collision.RobotModel(tm, bvh)
    q = np.array([[a, a, a] for a in expect])
    mask, _ = model.detect_batch(q)
    assert mask.cpu().numpy().tolist() == list(expect.values())
    rs = np.random.RandomState(2)
    qs = rs.uniform(-2.7, 2.7, size=(400, 3))
    mb, _ = model.detect_batch(qs)
    for b in range(0, 400, 57):
        for j, v in zip(model.joint_names, qs[b]):
            tm.set_joint(j, v)
        bvh.update_collider_poses()
        contacts = self_collision.detect(bvh)
        assert [int(contacts[f]) for f in frames] == mb[b].cpu().numpy().tolist()


def test_fused_aabb_and_candidate_filter_equals_the_two_calls():
    """d3d_aabb_filter_pairs = d3d_aabb + d3d_filter_pairs in one pass: same boxes, same pair set."""
    import ctypes
    import torch
    from distance3d_b200 import _lib
    from distance3d_b200._lib import c_i64, c_int, ptr
    tm = UrdfTransformManager()
    tm.load_urdf(open(os.path.join(DATA, "robot_arm.urdf")).read(), mesh_path=DATA)
    b = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
    b.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
    model = self_collision.RobotModel(tm, b)
    B = 3001                                          # not a multiple of the groups per block
    q = np.random.RandomState(8).uniform(-3.0, 3.0, size=(B, model.n_joints))
    dc = model.colliders_for(model.forward_kinematics(q))
    L = _lib.lib()
    n_pattern = int(model.pattern_t.shape[0])
    cap = B * n_pattern
    out = {}
    for fused in (False, True):
        aabb = torch.zeros((dc.n, 3, 2), dtype=torch.float64, device="cuda")
        pairs = torch.zeros((cap, 2), dtype=torch.int32, device="cuda")
        count = torch.zeros(1, dtype=torch.int64, device="cuda")
        if fused:
            _lib._check(L.d3d_aabb_filter_pairs(ctypes.byref(dc.struct), c_i64(B), c_int(model.n_frames),
                                                ptr(model.pattern_t), c_int(n_pattern), ptr(aabb), ptr(pairs),
                                                c_i64(cap), ptr(count), _lib.stream_ptr()))
        else:
            _lib._check(L.d3d_aabb(ctypes.byref(dc.struct), ptr(aabb), _lib.stream_ptr()))
            _lib._check(L.d3d_filter_pairs(ptr(aabb), c_i64(B), c_int(model.n_frames), ptr(model.pattern_t),
                                           c_int(n_pattern), ptr(pairs), c_i64(cap), ptr(count), _lib.stream_ptr()))
        n = int(count.item())
        p = pairs[:n].cpu().numpy()
        out[fused] = (aabb.cpu().numpy(), p[np.lexsort((p[:, 1], p[:, 0]))])
    assert np.array_equal(out[True][0], out[False][0])
    assert len(out[True][1]) > 1000 and np.array_equal(out[True][1], out[False][1])
