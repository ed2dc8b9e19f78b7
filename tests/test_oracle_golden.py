"""CPU: the C oracle reproduces the REAL reference's outputs bit for bit.

The fixtures in tests/golden/ were produced by oracle/gen_golden.py from the
reference itself (AlexanderFabisch/distance3d v0.9.0, numba JIT on).  This pins
the oracle; the GPU tests then compare the CUDA path with the oracle.
"""
import numpy as np

from oracle import cpu_oracle as O
from util import load_golden

MAX_FLOAT = np.finfo(float).max


def test_support_aabb_center_bit_exact():
    cs, g = load_golden("support.npz")
    O.prepare(cs)
    for i in range(len(cs)):
        for d, ref in zip(g["dirs"][i], g["support"][i]):
            np.testing.assert_array_equal(O.support(cs, i, d), ref)
        np.testing.assert_array_equal(O.center(cs, i), g["center"][i])
    np.testing.assert_array_equal(O.aabb(cs), g["aabb"])


def _check_gjk(cs, g):
    res = O.gjk_distance(cs, g["pairs"])
    ok = g["status"] != 4  # reference raised AssertionError (sanity check)
    assert np.all(res["status"][~ok] >= 4)
    clipped = g["status"] == 3
    assert np.all(res["status"][clipped] == 3)
    m = ok & ~clipped
    np.testing.assert_array_equal(res["dist"][m], g["dist"][m])
    np.testing.assert_array_equal(res["a"][m], g["a"][m])
    np.testing.assert_array_equal(res["b"][m], g["b"][m])
    np.testing.assert_array_equal(res["iters"][ok], g["iters"][ok])
    for k in np.where(m)[0]:
        n = res["n_points"][k]
        np.testing.assert_array_equal(res["Y"][k, :n], g["Y"][k, :n])
    hit = O.gjk_intersection(cs, g["pairs"])
    np.testing.assert_array_equal(hit["hit"][ok], g["hit"][ok])
    return res


def test_gjk_all_collider_types_bit_exact():
    cs, g = load_golden("gjk.npz")
    res = _check_gjk(cs, g)
    assert set(int(t) for t in np.unique(cs.type)) == set(range(10)) - {6}  # MeshGraph: test_meshgraph
    assert 0.05 < np.mean(res["dist"] == 0.0) < 0.6


def test_gjk_hulls_64_to_256_vertices_bit_exact():
    cs, g = load_golden("hulls.npz")
    _check_gjk(cs, g)
    assert cs.vert_len.min() >= 64 and cs.vert_len.max() <= 256


def test_epa_bit_exact_including_max_faces_assert():
    cs, g = load_golden("epa.npz")
    sel = np.where(g["status"] >= 0)[0]
    res = O.epa(cs, g["pairs"][sel], g["Y"][sel], return_faces=True)
    asserted = g["status"][sel] == 7
    np.testing.assert_array_equal(res["status"] == 7, asserted)
    m = ~asserted
    np.testing.assert_array_equal(res["mtv"][m], g["mtv"][sel][m])
    np.testing.assert_array_equal(res["success"][m], g["success"][sel][m])
    np.testing.assert_array_equal(res["n_faces"][m], g["n_faces"][sel][m])
    for q in np.where(m)[0]:
        n = res["n_faces"][q]
        np.testing.assert_array_equal(res["faces"][q, :n], g["faces"][sel[q], :n])
    assert asserted.sum() > 0 and m.sum() > 50


def test_epa_on_degenerate_simplices_bit_exact():
    """Real-reference outputs for simplices with a duplicated point or two points 1e-9 apart (closer
    than the edge-matching epsilon, epa.py:189-191)."""
    cs, g = load_golden("epa_degenerate.npz")
    res = O.epa(cs, g["pairs"], g["Y"], return_faces=True)
    asserted = g["status"] == 7
    np.testing.assert_array_equal(res["status"] == 7, asserted)
    m = ~asserted
    np.testing.assert_array_equal(res["mtv"][m], g["mtv"][m])
    np.testing.assert_array_equal(res["success"][m], g["success"][m])
    np.testing.assert_array_equal(res["n_faces"][m], g["n_faces"][m])
    for q in np.where(m)[0]:
        n = res["n_faces"][q]
        np.testing.assert_array_equal(res["faces"][q, :n], g["faces"][q, :n])
    for kind in (0, 1, 2):
        assert (m & (g["kind"] == kind)).sum() > 100


def test_epa_on_wide_hulls_bit_exact():
    cs, g = load_golden("hulls.npz")
    sel = np.where(g["epa_status"] >= 0)[0]
    res = O.epa(cs, g["pairs"][sel], g["Y"][sel])
    asserted = g["epa_status"][sel] == 7
    np.testing.assert_array_equal(res["status"] == 7, asserted)
    np.testing.assert_array_equal(res["mtv"][~asserted], g["epa_mtv"][sel][~asserted])


def test_mpr_bit_exact():
    cs, g = load_golden("mpr.npz")
    res = O.mpr(cs, g["pairs"])
    np.testing.assert_array_equal(res["hit"], g["hit"])
    hit = g["hit"].astype(bool)
    np.testing.assert_array_equal(res["depth"][hit], g["depth"][hit])
    np.testing.assert_array_equal(res["dir"][hit], g["dir"][hit])
    np.testing.assert_array_equal(res["pos"][hit], g["pos"][hit])
    res_i = O.mpr(cs, g["pairs"], penetration=False)
    np.testing.assert_array_equal(res_i["hit"], g["hit_intersection"])


def test_broad_phase_tree_and_brute_force_identical_lists():
    cs, g = load_golden("aabb.npz")
    A = O.aabb(cs)
    np.testing.assert_array_equal(A, g["aabb"])
    tree = O.Tree()
    tree.insert_aabbs(A)
    np.testing.assert_array_equal(tree.nodes, g["tree_nodes"])
    assert tree.root == int(g["tree_root"])
    np.testing.assert_array_equal(tree.overlaps_aabb_tree(tree), g["tree_pairs"])
    np.testing.assert_array_equal(O.all_aabbs_overlap(A[:300], A[300:700]), g["brute_pairs"])
    # tree result == brute force as a set
    brute = O.all_aabbs_overlap(A, A)
    assert set(map(tuple, brute)) == set(map(tuple, g["tree_pairs"]))


def test_self_collision_masks_equal_reference_detect():
    """FK + AABB + white-list + gjk_intersection of the oracle reproduce the reference's
    self_collision.detect over 150 joint configurations (tests/golden/self_collision.npz)."""
    import os
    from distance3d_b200 import colliders as C, pack
    from distance3d_b200.urdf import UrdfTransformManager
    from distance3d_b200.urdf_utils import self_collision_whitelists
    from distance3d_b200.self_collision import _candidate_pattern
    from util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "self_collision.npz"))
    data = os.path.join(os.path.dirname(GOLDEN), "data")
    tm = UrdfTransformManager()
    with open(os.path.join(data, "robot_arm.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=data)
    tm.add_transform("robot_arm", "origin", np.eye(4))
    frames = [str(f) for f in g["frames"]]
    template = pack.pack_colliders([C.Cylinder(np.eye(4), o.radius, o.length) for o in tm.collision_objects])
    pattern, wl, symmetric = _candidate_pattern(frames, self_collision_whitelists(tm))
    assert symmetric and len(pattern) == 17
    kin = tm.compile_kinematics(frames, "origin")
    poses = O.fk(kin, g["q"])
    assert np.max(np.abs(poses - g["poses"])) < 1e-14
    masks, n_cand = O.self_collision_masks(template, kin, pattern, g["q"])
    np.testing.assert_array_equal(masks, g["mask"])
    assert n_cand > 100
    wl_matrix = (wl[:, None] >> np.arange(len(frames), dtype=np.uint64)[None, :]) & np.uint64(1)
    ordered = O.self_collision_masks_ordered(template, kin, wl_matrix, g["q"][:40])
    np.testing.assert_array_equal(ordered, g["mask"][:40])  # order cannot matter for a chain


def _branched_robot():
    import os
    from distance3d_b200 import colliders as C, pack
    from distance3d_b200.urdf import UrdfTransformManager, Cylinder
    from util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "self_collision_branched.npz"))
    data = os.path.join(os.path.dirname(GOLDEN), "data")
    tm = UrdfTransformManager()
    with open(os.path.join(data, "robot_branched.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=data)
    tm.add_transform("robot_branched", "origin", np.eye(4))
    template = pack.pack_colliders([
        C.Cylinder(np.eye(4), o.radius, o.length) if isinstance(o, Cylinder)
        else C.Sphere(np.zeros(3), o.radius) for o in tm.collision_objects])
    return g, tm, template


def test_branched_robot_whitelists_and_ordered_detect_equal_reference():
    """A link with several children white-lists only the last one (urdf_utils.py:79-81): the
    white-lists are asymmetric and the reference's detect() depends on the order of its
    AABB-tree candidates.  The fixture holds the reference's white-lists and masks."""
    from distance3d_b200.urdf_utils import self_collision_whitelists
    from distance3d_b200.self_collision import _candidate_pattern
    g, tm, template = _branched_robot()
    frames = [str(f) for f in g["frames"]]
    assert [o.frame for o in tm.collision_objects] == frames
    wl = self_collision_whitelists(tm)
    wl_matrix = np.array([[int(fj in wl[fi]) for fj in frames] for fi in frames], dtype=np.uint8)
    np.testing.assert_array_equal(wl_matrix, g["wl"])
    pattern, bits, symmetric = _candidate_pattern(frames, wl)
    assert not symmetric
    assert (0, 1) in map(tuple, pattern)  # torso vs left shoulder: white-listed one way only
    kin = tm.compile_kinematics(frames, "origin")
    assert list(kin["joint_names"]) == [str(j) for j in g["joints"]]
    masks = O.self_collision_masks_ordered(template, kin, wl_matrix, g["q"])
    np.testing.assert_array_equal(masks, g["mask"])
    np.testing.assert_array_equal(masks.any(axis=1), g["any"].astype(bool))


def test_meshgraph_hill_climbing_bit_exact():
    """MeshGraph support = hill climbing from a cached vertex (mesh.py:12-139).  The fixture
    was generated with a fresh reference object per call, so it is order independent."""
    cs, g = load_golden("meshgraph.npz")
    assert cs.graph_off is not None and (cs.graph_off[cs.type == 6] >= 0).all()
    # the adjacency records of the fixture are what build_mesh_graph makes of the triangles
    from distance3d_b200.mesh import build_mesh_graph
    t0 = 0
    for m, nt in enumerate(g["tri_len"]):
        tri = g["triangles"][t0:t0 + nt]
        t0 += nt
        o, l = cs.vert_off[m], cs.vert_len[m]
        rec = build_mesh_graph(cs.verts[o:o + l], tri)
        np.testing.assert_array_equal(rec, cs.graph[cs.graph_off[m]:cs.graph_off[m] + len(rec)])
    res = _check_gjk(cs, g)
    assert (res["dist"] == 0.0).sum() > 100 and (res["dist"] > 0.0).sum() > 100
    # EPA
    sel = np.where(g["epa_status"] >= 0)[0]
    e = O.epa(cs, g["pairs"][sel], g["Y"][sel])
    asserted = g["epa_status"][sel] == 7
    np.testing.assert_array_equal(e["status"] == 7, asserted)
    np.testing.assert_array_equal(e["mtv"][~asserted], g["epa_mtv"][sel][~asserted])
    np.testing.assert_array_equal(e["success"][~asserted], g["epa_success"][sel][~asserted])
    np.testing.assert_array_equal(e["n_faces"][~asserted], g["epa_n_faces"][sel][~asserted])
    # MPR
    m = O.mpr(cs, g["pairs"], penetration=True)
    np.testing.assert_array_equal(m["hit"], g["mpr_hit"])
    h = m["hit"].astype(bool)
    np.testing.assert_array_equal(m["depth"][h], g["mpr_depth"][h])
    np.testing.assert_array_equal(m["dir"][h], g["mpr_dir"][h])
    np.testing.assert_array_equal(m["pos"][h], g["mpr_pos"][h])
    np.testing.assert_array_equal(O.mpr(cs, g["pairs"], penetration=False)["hit"], g["mpr_hit_intersection"])
    # one object, consecutive support calls: the vertex is carried from call to call
    for mi in range(len(g["seq_dirs"])):
        start = None
        for t in range(g["seq_dirs"].shape[1]):
            pt, start = O.support(cs, mi, g["seq_dirs"][mi, t], start=start, return_index=True)
            np.testing.assert_array_equal(pt, g["seq_pts"][mi, t])
            assert start == g["seq_idx"][mi, t]


def test_tetrahedron_pairs_vs_reference_outputs():
    """Hydroelastic broad-phase consumer: the oracle's intersect_tetrahedron_pairs
    against the reference's on 15 pairs of tetrahedral meshes made by the reference's generators
    (270 k candidate pairs from all_aabbs_overlap).  Booleans identical, planes and polygon
    regions within 1e-9 (tests/util.py compare_tetra_results)."""
    import os
    from util import GOLDEN, compare_tetra_results
    g = dict(np.load(os.path.join(GOLDEN, "tetra.npz")))
    compared = 0
    for c in range(int(g["n_cases"])):
        k = "c%d_" % c
        np.testing.assert_array_equal(O.tetra_aabbs(g[k + "tp1"]), g[k + "aabb1"])
        np.testing.assert_array_equal(O.tetra_aabbs(g[k + "tp2"]), g[k + "aabb2"])
        # the candidate list of the fixture is the brute-force overlap of the boxes
        np.testing.assert_array_equal(O.all_aabbs_overlap(g[k + "aabb1"], g[k + "aabb2"]), g[k + "pairs"])
        r = O.tetra_pairs(g[k + "pairs"], g[k + "tp1"], g[k + "e1"], g[k + "tp2"], g[k + "e2"],
                          youngs_modulus1=g[k + "ym"][0], youngs_modulus2=g[k + "ym"][1],
                          n_threads=O.max_threads())
        compared += compare_tetra_results(r, g, k)
    assert compared > 2500


def test_gjk_intersection_libccd_vs_reference_outputs():
    """The reference's second boolean GJK (gjk/_gjk_libccd.py:14-266), all collider types +
    Margin + MeshGraph: identical booleans on 5000 pairs; and, as in the reference's own
    test_gjk.py:341-354, agreement with the Jolt variant."""
    for tag in ("far", "near"):
        cs, g = load_golden("libccd.npz", prefix=tag + "_cs_")
        got = O.gjk_intersection_libccd(cs, g[tag + "_pairs"])
        np.testing.assert_array_equal(got["hit"], g[tag + "_hit"])
        np.testing.assert_array_equal(O.gjk_intersection(cs, g[tag + "_pairs"])["hit"], g[tag + "_hit_jolt"])
        assert 0.1 < g[tag + "_hit"].mean() < 0.8
