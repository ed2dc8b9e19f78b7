"""CPU: host-side logic (packing, URDF kinematics, white-lists, random generators) and
the C ABI of the CUDA library (loads and exports every symbol of include/d3d_b200.h)."""
import ctypes
import os
import re

import numpy as np
import pytest

from distance3d_b200 import colliders as C, pack, random as d3random
from distance3d_b200.urdf import UrdfTransformManager, TransformManager
from distance3d_b200.urdf_utils import self_collision_whitelists, fast_transform_manager_initialization
from util import GOLDEN

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(REPO, "tests", "data")


def test_library_exports_every_declared_symbol():
    so = os.path.join(REPO, "distance3d_b200", "libd3d_b200.so")
    if not os.path.exists(so):
        from distance3d_b200 import build
        build.build()
    lib = ctypes.CDLL(so)
    header = open(os.path.join(REPO, "include", "d3d_b200.h")).read()
    names = set(re.findall(r"\b(d3d_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert lib.d3d_last_error_string is not None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "distance3d_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "libd3d_oracle" not in text and "cpu_oracle" not in text, f


def test_missing_library_fails_loudly(monkeypatch):
    from distance3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libd3d_b200.so")
    with pytest.raises(_lib.D3DError):
        _lib.lib()


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from distance3d_b200 import gjk, _lib
    with pytest.raises(_lib.D3DError):
        gjk.gjk(C.Sphere(np.zeros(3), 1.0), C.Sphere(np.ones(3), 1.0))


def test_pack_all_collider_types_and_margin():
    T = np.eye(4)
    cols = [C.Sphere(np.array([1.0, 2, 3]), 0.5), C.Capsule(T, 0.1, 0.2), C.Box(T, np.array([1.0, 2, 3])),
            C.Ellipsoid(T, np.array([1.0, 2, 3])), C.Cylinder(T, 0.3, 0.4),
            C.ConvexHullVertices(np.arange(12.0).reshape(4, 3)),
            C.MeshGraph(T, np.arange(15.0).reshape(5, 3), np.array([[0, 1, 2]])),
            C.Disk(np.zeros(3), 1.0, np.array([0.0, 0, 1])),
            C.Ellipse(np.zeros(3), np.eye(3)[:2], np.array([1.0, 2])), C.Cone(T, 1.0, 2.0),
            C.Margin(C.Sphere(np.zeros(3), 1.0), 0.25)]
    cs = pack.pack_colliders(cols)
    assert list(cs.type[:10]) == list(range(10)) and cs.type[10] == pack.SPHERE
    assert cs.margin[10] == 0.25 and cs.margin[:10].sum() == 0
    assert list(cs.vert_len) == [0, 0, 8, 0, 0, 4, 5, 0, 0, 0, 0]
    assert cs.vert_off[5] == 8 and cs.vert_off[6] == 12 and cs.n_vertices == 17
    np.testing.assert_array_equal(cs.pose[0, :3, 3], [1, 2, 3])
    sub = cs.subset([5, 2])
    assert list(sub.vert_len) == [4, 8] and sub.n_vertices == 12
    np.testing.assert_array_equal(sub.verts[:4], np.arange(12.0).reshape(4, 3))
    assert set(C.COLLIDERS) == {"sphere", "ellipsoid", "capsule", "disk", "ellipse", "cone",
                                "cylinder", "box", "mesh"}


def test_random_collider_set_shapes():
    rs = np.random.RandomState(0)
    cs = d3random.random_collider_set(rs, 5000, names=d3random.PRIMITIVES + ("mesh",),
                                      hull_vertices=(6, 12))
    assert len(cs) == 5000 and set(np.unique(cs.type)) == {0, 1, 2, 3, 4, 5}
    R = cs.pose[:, :3, :3]
    np.testing.assert_allclose(np.einsum("nij,nkj->nik", R, R), np.tile(np.eye(3), (5000, 1, 1)), atol=1e-12)
    hull = cs.type == pack.HULL
    assert cs.vert_len[hull].min() >= 6 and cs.vert_len[hull].max() <= 12
    assert cs.vert_off[-1] + cs.vert_len[-1] == cs.n_vertices
    pairs = d3random.random_pairs(rs, 5000, 1000)
    assert np.all(pairs[:, 0] != pairs[:, 1])
    for name, gen in d3random.RANDOM_GENERATORS.items():
        args = gen(np.random.RandomState(1))
        assert C.COLLIDERS[name](*args) is not None


def load_robot_tm():
    tm = UrdfTransformManager()
    with open(os.path.join(DATA, "robot_arm.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=DATA)
    return tm


def test_urdf_kinematics_match_reference_poses():
    g = np.load(os.path.join(GOLDEN, "self_collision.npz"))
    tm = load_robot_tm()
    tm.add_transform("robot_arm", "origin", np.eye(4))
    frames = [str(f) for f in g["frames"]]
    assert [o.frame for o in tm.collision_objects] == frames
    kin = tm.compile_kinematics(frames, "origin")
    assert kin["joint_names"] == ["joint%d" % i for i in range(1, 7)]
    for b in (0, 2, 17):
        for j in range(6):
            tm.set_joint("joint%d" % (j + 1), g["q"][b, j])
        for k, fr in enumerate(frames):
            np.testing.assert_allclose(tm.get_transform(fr, "origin"), g["poses"][b, k], atol=1e-14)
            # flattened chain reproduces the graph walk
            T = np.eye(4)
            for s in range(kin["chain_off"][k], kin["chain_off"][k + 1]):
                T = T @ kin["chain_fixed"][s]
                j = kin["chain_joint"][s]
                if j >= 0:
                    from distance3d_b200._transforms import matrix_from_axis_angle, transform_from
                    T = T @ transform_from(matrix_from_axis_angle(kin["joint_axis"][j], g["q"][b, j]), np.zeros(3))
            np.testing.assert_allclose(T, g["poses"][b, k], atol=1e-13)


def test_kinematic_tree_is_the_chains_with_shared_prefixes():
    """compile_kinematics also emits the chain steps as a tree (d3d_fk_urdf_tree): walking from a
    frame's node to the root gives back exactly that frame's chain, parents come first, and the arm
    needs one node per joint plus one tail per frame."""
    for urdf in ("robot_arm.urdf", "robot_branched.urdf"):
        tm = UrdfTransformManager()
        with open(os.path.join(DATA, urdf)) as f:
            tm.load_urdf(f.read(), mesh_path=DATA)
        tm.add_transform(urdf[:-5], "origin", np.eye(4))
        frames = [o.frame for o in tm.collision_objects]
        kin = tm.compile_kinematics(frames, "origin")
        n = len(kin["node_parent"])
        assert np.all(kin["node_parent"] < np.arange(n))
        end = np.empty(len(frames), dtype=int)
        for node in range(n):
            end[kin["node_out"][kin["node_out_off"][node]:kin["node_out_off"][node + 1]]] = node
        steps_flat = steps_tree = 0
        for k in range(len(frames)):
            path, node = [], end[k]
            while node >= 0:
                path.append(node)
                node = kin["node_parent"][node]
            path = path[::-1]
            lo, hi = kin["chain_off"][k], kin["chain_off"][k + 1]
            assert len(path) == hi - lo
            for node, s_ in zip(path, range(lo, hi)):
                assert np.array_equal(kin["node_fixed"][node], kin["chain_fixed"][s_])
                assert kin["node_joint"][node] == kin["chain_joint"][s_]
            steps_flat += hi - lo
        has_child = np.zeros(n, dtype=bool)
        has_child[kin["node_parent"][kin["node_parent"] >= 0]] = True
        assert np.array_equal(kin["node_keep"] >= 0, has_child) and kin["n_keep"] == has_child.sum()
        assert sorted(kin["node_keep"][has_child]) == list(range(kin["n_keep"]))
        assert n < steps_flat          # prefixes are shared
        if urdf == "robot_arm.urdf":
            assert n == 6 + len(frames) and kin["n_keep"] == 6


def test_self_collision_whitelists_match_survey():
    tm = load_robot_tm()
    wl = self_collision_whitelists(tm)
    short = {k.split(":")[1]: sorted(f.split(":")[1] for f in v) for k, v in wl.items()}
    assert short["link1/0"] == ["link1/0", "link2/0", "link2/1"]
    assert short["link4/0"] == ["link3/0", "link4/0", "link5/0", "link5/1"]
    assert short["link6/0"] == ["link5/0", "link5/1", "link6/0"]
    tm2 = TransformManager()
    fast_transform_manager_initialization(tm2, [1, 2, 3], "base")
    fast_transform_manager_initialization(tm2, [4, 5], 1)
    np.testing.assert_allclose(tm2.get_transform(5, "base"), np.eye(4))


def test_c_abi_argument_checks_without_a_gpu():
    """Misuse is reported through return codes + d3d_last_error_string before any CUDA call."""
    so = os.path.join(REPO, "distance3d_b200", "libd3d_b200.so")
    lib = ctypes.CDLL(so)
    lib.d3d_last_error_string.restype = ctypes.c_char_p
    lib.d3d_gjk_workspace_bytes.restype = ctypes.c_size_t
    lib.d3d_bvh_workspace_bytes.restype = ctypes.c_size_t
    lib.d3d_bvh_query_workspace_bytes.restype = ctypes.c_size_t
    i64, dbl, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    # workspace sizes grow with the problem and never return 0
    assert 0 < lib.d3d_gjk_workspace_bytes(i64(0)) < lib.d3d_gjk_workspace_bytes(i64(1 << 20))
    assert 0 < lib.d3d_bvh_workspace_bytes(i64(1)) < lib.d3d_bvh_workspace_bytes(i64(1 << 20))
    assert lib.d3d_bvh_query_workspace_bytes(i64(1 << 20)) >= (1 << 20) * 12
    # empty batches are a no-op
    assert lib.d3d_gjk_distance(None, None, i64(0), dbl(1e-10), dbl(1e5), dbl(1e-8), None, None, None,
                                None, None, None, None, None, ctypes.c_size_t(0), None) == 0
    assert lib.d3d_epa(None, None, i64(0), None, None, 64, 32, 64, dbl(1e-8), None, None, None, None, None,
                       None, None, ctypes.c_size_t(0), None) == 0
    # null arguments / bad limits
    assert lib.d3d_gjk_distance(None, None, i64(5), dbl(1e-10), dbl(1e5), dbl(1e-8), None, None, None,
                                None, None, None, None, None, ctypes.c_size_t(0), None) == -1
    assert b"null argument" in lib.d3d_last_error_string()
    cs = pack.pack_colliders([C.Sphere(np.zeros(3), 1.0)]).host_struct()
    dummy = (ctypes.c_double * 64)()
    pairs = (ctypes.c_int32 * 2)(0, 0)
    rc = lib.d3d_epa(ctypes.byref(cs), pairs, i64(1), dummy, None, 64, 32, 65, dbl(1e-8), dummy,
                     ctypes.cast(dummy, vp), None, None, None, None, ctypes.cast(dummy, vp),
                     ctypes.c_size_t(1 << 20), None)
    assert rc == -1 and b"max_faces" in lib.d3d_last_error_string()
    lib.d3d_epa_workspace_bytes.restype = ctypes.c_size_t
    assert lib.d3d_epa_workspace_bytes(i64(1 << 20)) >= 5 << 20   # processing order: 5 B per pair
    rc = lib.d3d_epa(ctypes.byref(cs), pairs, i64(1), dummy, None, 64, 32, 64, dbl(1e-8), dummy,
                     ctypes.cast(dummy, vp), None, None, None, None, ctypes.cast(dummy, vp),
                     ctypes.c_size_t(64), None)
    assert rc == -1 and b"workspace too small" in lib.d3d_last_error_string()
    assert lib.d3d_bvh_build(None, i64(-1), ctypes.cast(dummy, vp), ctypes.c_size_t(512), None) == -1
    assert b"out of range" in lib.d3d_last_error_string()
    assert lib.d3d_bvh_build(dummy, i64(1000), ctypes.cast(dummy, vp), ctypes.c_size_t(512), None) == -1
    assert b"workspace too small" in lib.d3d_last_error_string()


def test_bench_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` times the CPU oracle only and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--pairs", "20000", "--cpu-sample", "20000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gjk_distance_pairs_per_s"
    assert line["value"] > 0 and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


MESH_URDF = """<?xml version="1.0"?>
<robot name="mesh_bot">
  <link name="base"><collision><origin xyz="0 0 0.1" rpy="0 0 0"/>
    <geometry><mesh filename="part.stl" scale="1 2 0.5"/></geometry></collision></link>
  <link name="arm"><collision><origin xyz="0 0 0" rpy="0 0 0"/>
    <geometry><mesh filename="part.stl"/></geometry></collision></link>
  <link name="tip"><collision><geometry><mesh filename="missing.stl"/></geometry></collision></link>
  <joint name="j1" type="revolute"><parent link="base"/><child link="arm"/>
    <origin xyz="0 0 1" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
  <joint name="j2" type="revolute"><parent link="arm"/><child link="tip"/>
    <origin xyz="0 0 1" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
</robot>"""


def test_mesh_io_round_trips_and_urdf_mesh_colliders(tmp_path):
    """distance3d/io.py:5-46 load_mesh + broad_phase.py:93-127 _make_collider for <mesh>."""
    import warnings
    from distance3d_b200 import io, mesh, broad_phase
    from distance3d_b200.urdf import UrdfTransformManager
    rs = np.random.RandomState(5)
    V = rs.randn(30, 3)
    V -= V.mean(axis=0)
    T = mesh.make_convex_mesh(V)
    used = np.unique(T)
    for binary, expect in ((True, V.astype(np.float32).astype(np.float64)), (False, V)):
        f = str(tmp_path / ("b.stl" if binary else "a.stl"))
        io.save_stl(f, V, T, binary=binary)
        V2, T2 = io.load_mesh(f, scale=2.0)
        assert len(V2) == len(used) and T2.shape == T.shape          # duplicated corners are merged
        np.testing.assert_array_equal(V2[T2], 2.0 * expect[T])         # same triangles, same order
    obj = tmp_path / "m.obj"
    obj.write_text("# quad\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1/1 2/2 3/3 4/4\nf -4 -3 -2\n")
    Vo, To = io.load_mesh(str(obj))
    assert Vo.shape == (4, 3) and To.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    with pytest.raises(OSError):
        io.load_mesh(str(tmp_path / "x.ply"))
    # URDF <mesh> -> MeshGraph in the collision frame, scaled per axis
    io.save_stl(str(tmp_path / "part.stl"), V, T, binary=False)
    tm = UrdfTransformManager()
    tm.load_urdf(MESH_URDF, mesh_path=str(tmp_path))
    bvh = broad_phase.BoundingVolumeHierarchy.__new__(broad_phase.BoundingVolumeHierarchy)
    tm.add_transform("mesh_bot", "origin", np.eye(4))
    made = {}
    for o in tm.collision_objects:
        try:
            made[o.frame] = bvh._make_collider(tm, o, False)
        except RuntimeError as e:                     # the missing file surfaces like in the reference
            assert "missing.stl" in str(e)
            made[o.frame] = None
    frames = sorted(made)
    assert len(frames) == 3 and sum(v is None for v in made.values()) == 1
    base = made["collision:base/0"]
    assert isinstance(base, C.MeshGraph)
    np.testing.assert_array_equal(base.vertices[base.triangles], (V * [1.0, 2.0, 0.5])[T])
    np.testing.assert_allclose(base.mesh2origin[:3, 3], [0.0, 0.0, 0.1])
    arm = made["collision:arm/0"]
    np.testing.assert_allclose(arm.mesh2origin[:3, 3], [0.0, 0.0, 1.0])
    packed = pack.pack_colliders([base, arm])
    assert packed.type.tolist() == [pack.MESH, pack.MESH] and packed.vert_len.tolist() == [len(used)] * 2


def test_host_wire_packer_records_and_sizes():
    """d3d_wire_size / d3d_pack_wire_host are host code (no GPU): offsets are the exclusive scan of
    the record sizes, the records hold the pose / parameter / vertex-range fields of their type, and
    the result does not depend on the number of packing threads."""
    import ctypes
    from distance3d_b200 import _lib, colliders as C, pack as P, random as R
    rs = np.random.RandomState(4)
    cs = R.random_collider_set(rs, 3000, names=R.PRIMITIVES + ("mesh", "cone"), hull_vertices=(4, 12))
    extra = P.pack_colliders([C.Disk(rs.randn(3), 0.7, np.array([0.0, 0.6, 0.8])),
                              C.Ellipse(rs.randn(3), np.eye(3)[:2], np.array([0.4, 0.9]))])
    cs = P.concat_sets([cs, extra])
    sizes = np.array([4, 14, 16, 15, 14, 1, 13, 7, 11, 14])[cs.type]
    wt, wo, w = cs.wire(n_threads=1)
    assert np.array_equal(wt, cs.type.astype(np.uint8))
    assert np.array_equal(wo, np.concatenate(([0], np.cumsum(sizes)[:-1])))
    assert len(w) == sizes.sum()
    wt8, wo8, w8 = cs.wire(n_threads=8)
    assert np.array_equal(wt8, wt) and np.array_equal(wo8, wo) and np.array_equal(w8.view(np.uint64), w.view(np.uint64))
    for i in range(len(cs)):
        r = w[wo[i]:wo[i] + sizes[i]]
        t, T, p = cs.type[i], cs.pose[i], cs.param[i]
        rng = np.array([cs.vert_off[i], cs.vert_len[i]], dtype=np.int32).view(np.float64)[0]
        if t == P.SPHERE:
            expect = [T[0, 3], T[1, 3], T[2, 3], p[0]]
        elif t in (P.CAPSULE, P.CYLINDER, P.CONE):
            expect = list(T[:3].ravel()) + [p[0], p[1]]
        elif t == P.ELLIPSOID:
            expect = list(T[:3].ravel()) + list(p)
        elif t == P.BOX:
            expect = list(T[:3].ravel()) + list(p) + [rng]
        elif t == P.HULL:
            expect = [rng]
        elif t == P.DISK:
            expect = [T[0, 3], T[1, 3], T[2, 3], T[0, 2], T[1, 2], T[2, 2], p[0]]
        else:   # ellipse
            expect = [T[0, 3], T[1, 3], T[2, 3], T[0, 0], T[1, 0], T[2, 0], T[0, 1], T[1, 1], T[2, 1], p[0], p[1]]
        assert np.array_equal(np.asarray(expect).view(np.uint64), r.view(np.uint64)), (i, t)
    # the size pass is shared by the host threads above 256 k colliders
    big = np.ascontiguousarray(rs.randint(0, 10, size=1_300_001), dtype=np.int32)
    L = _lib.lib()
    L.d3d_wire_size.restype = ctypes.c_int64
    assert L.d3d_wire_size(ctypes.c_void_p(big.ctypes.data), ctypes.c_int64(len(big))) == \
        int(np.array([4, 14, 16, 15, 14, 1, 13, 7, 11, 14])[big].sum())
    big[777_777] = 12
    assert L.d3d_wire_size(ctypes.c_void_p(big.ctypes.data), ctypes.c_int64(len(big))) == -1
    assert b"unknown collider type 12 at 777777" in L.d3d_last_error_string()
