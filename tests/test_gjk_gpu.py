"""GPU: the CUDA GJK path (through the C ABI) against the CPU oracle.

Tolerances are the ones BASELINE.json states: distance and closest points within
1e-9 absolute; intersection booleans identical except for pairs whose oracle
distance is within 1e-12 of contact.  Bit-exact agreement is additionally
reported (and required for the collider types whose support maps contain no
BLAS-norm call: capsule, cylinder, box, hull).
"""
import numpy as np
import pytest

from distance3d_b200 import gjk, random as d3random
from distance3d_b200 import pack as P
from oracle import cpu_oracle as O
from util import load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


def compare_distance(cs, pairs, exact_types=None, **kw):
    res = gjk.gjk_distance_batch(cs, pairs, **kw).cpu()
    ref = O.gjk_distance(cs, pairs, n_threads=O.max_threads(), **kw)
    ok = ref["status"] <= 1
    # identical termination state, except within 1e-12 of contact
    near = (ref["dist"] < 1e-12) | (np.abs(res["dist"]) < 1e-12)
    same = res["status"] == ref["status"]
    assert np.all(same | near), "status differs on %d pairs" % int((~(same | near)).sum())
    clipped = ref["status"] == 3
    assert np.array_equal(res["status"][clipped], ref["status"][clipped])
    m = ok & same
    assert np.max(np.abs(res["dist"][m] - ref["dist"][m]), initial=0.0) < TOL
    assert np.max(np.abs(res["closest_a"][m] - ref["a"][m]), initial=0.0) < TOL
    assert np.max(np.abs(res["closest_b"][m] - ref["b"][m]), initial=0.0) < TOL
    if exact_types is not None:
        pairs = np.asarray(pairs)
        ex = np.isin(cs.type[pairs[:, 0]], exact_types) & np.isin(cs.type[pairs[:, 1]], exact_types)
        if True:
            e = ex & ok
            assert np.array_equal(res["dist"][e], ref["dist"][e])
            assert np.array_equal(res["closest_a"][e], ref["a"][e])
            assert np.array_equal(res["closest_b"][e], ref["b"][e])
            assert np.array_equal(res["iters"][ex], ref["iters"][ex])
            assert np.array_equal(res["n_points"][e], ref["n_points"][e])
            for k in np.where(e)[0][:500]:
                n = ref["n_points"][k]
                assert np.array_equal(res["simplex"][k, :n], ref["Y"][k, :n])
    return res, ref


EXACT = tuple(range(10))  # every collider type is bit-exact (x87 norm emulated)


def test_golden_all_types_vs_reference_outputs():
    """CUDA path against the outputs of the real reference stored in tests/golden."""
    cs, g = load_golden("gjk.npz")
    res = gjk.gjk_distance_batch(cs, g["pairs"]).cpu()
    ok = g["status"] <= 1
    assert np.max(np.abs(res["dist"][ok] - g["dist"][ok])) < TOL
    assert np.max(np.abs(res["closest_a"][ok] - g["a"][ok])) < TOL
    assert np.max(np.abs(res["closest_b"][ok] - g["b"][ok])) < TOL
    assert np.array_equal(res["status"][g["status"] == 3], g["status"][g["status"] == 3])
    hit, _, _ = gjk.gjk_intersection_batch(cs, g["pairs"])
    hit = hit.cpu().numpy()
    near = g["dist"] < 1e-12
    valid = g["status"] != 4
    assert np.array_equal(hit[valid & ~near], g["hit"][valid & ~near])
    compare_distance(cs, g["pairs"], exact_types=EXACT)
    # bit-exact against the real reference's outputs, all nine collider types + Margin
    assert np.array_equal(res["dist"][ok], g["dist"][ok])
    assert np.array_equal(res["closest_a"][ok], g["a"][ok])
    assert np.array_equal(res["closest_b"][ok], g["b"][ok])
    valid = g["status"] != 4
    assert np.array_equal(res["iters"][valid], g["iters"][valid])
    assert np.array_equal(hit[valid], g["hit"][valid])


def test_golden_wide_hulls_warp_kernel():
    cs, g = load_golden("hulls.npz")
    res, ref = compare_distance(cs, g["pairs"], exact_types=EXACT)
    assert np.array_equal(res["dist"], g["dist"])  # bit-exact against the real reference
    assert np.array_equal(res["closest_a"], g["a"])


@pytest.mark.parametrize("names,exact", [
    (("capsule", "cylinder", "box"), True),
    (d3random.PRIMITIVES, True),
    (d3random.PRIMITIVES + ("mesh",), True),
    (("cone", "sphere", "box"), True),
])
def test_random_batches(names, exact):
    rs = np.random.RandomState(11)
    cs = d3random.random_collider_set(rs, 3000, names=names)
    pairs = d3random.random_pairs(rs, len(cs), 40000)
    compare_distance(cs, pairs, exact_types=EXACT if exact else None)


def test_mixed_wide_and_small_hulls():
    rs = np.random.RandomState(12)
    cs = d3random.random_collider_set(rs, 600, names=("mesh", "box", "capsule"),
                                      hull_vertices=(4, 120), center_scale=1.5)
    pairs = d3random.random_pairs(rs, len(cs), 6000)
    compare_distance(cs, pairs, exact_types=EXACT)


def test_intersection_matches_oracle():
    rs = np.random.RandomState(13)
    cs = d3random.random_collider_set(rs, 3000, names=d3random.PRIMITIVES + ("mesh",))
    pairs = d3random.random_pairs(rs, len(cs), 50000)
    hit, iters, status = gjk.gjk_intersection_batch(cs, pairs, want_iters=True)
    ref = O.gjk_intersection(cs, pairs, n_threads=O.max_threads())
    dist = O.gjk_distance(cs, pairs, n_threads=O.max_threads())["dist"]
    hit = hit.cpu().numpy()
    assert np.array_equal(hit, ref["hit"])
    assert np.array_equal(iters.cpu().numpy(), ref["iters"])
    assert np.array_equal(status.cpu().numpy(), ref["status"])
    assert np.array_equal(hit == 1, dist == 0.0)
    assert 0.05 < hit.mean() < 0.6


def test_edge_cases_empty_single_clipped_identical():
    rs = np.random.RandomState(14)
    cs = d3random.random_collider_set(rs, 50, names=d3random.PRIMITIVES)
    res = gjk.gjk_distance_batch(cs, np.zeros((0, 2), dtype=np.int32)).cpu()
    assert res["dist"].shape == (0,)
    # a collider against itself intersects
    pairs = np.stack([np.arange(50), np.arange(50)], axis=1).astype(np.int32)
    res, ref = compare_distance(cs, pairs)
    assert np.all(res["dist"] == 0.0)
    # far apart -> clipped (MAX_FLOAT)
    cs.pose[:25, :3, 3] += 1.0e4
    cs.invalidate_device()
    pairs = np.stack([np.arange(25), np.arange(25, 50)], axis=1).astype(np.int32)
    res, ref = compare_distance(cs, pairs)
    assert np.all(res["status"] == 3) and np.all(res["dist"] == np.finfo(float).max)


def test_scalar_api_matches_reference_semantics():
    from distance3d_b200 import colliders as C
    s1 = C.Sphere(np.zeros(3), 1.0)
    s2 = C.Sphere(np.array([0.0, 0.0, 3.0]), 1.0)
    d, a, b, Y = gjk.gjk(s1, s2)
    assert abs(d - 1.0) < 1e-9 and Y.shape == (4, 3)
    assert gjk.gjk(s1, C.Sphere(np.array([0.0, 0.0, 1e3]), 1.0)) == (np.finfo(float).max, None, None, None)
    assert gjk.gjk_intersection(s1, C.Sphere(np.array([0.0, 0.5, 0.0]), 1.0)) is True
    assert gjk.gjk_intersection(s1, s2) is False
    assert gjk.gjk_distance_jolt_iterations(s1, s2) >= 1
    box = C.Box(np.eye(4), np.ones(3))
    np.testing.assert_allclose(box.aabb(), [[-0.5, 0.5]] * 3)
    np.testing.assert_allclose(box.support_function(np.array([1.0, 1.0, 1.0])), [0.5, 0.5, 0.5])


def test_meshgraph_hill_climbing_bit_exact_vs_reference_outputs(monkeypatch):
    """MeshGraph support = hill climbing over the triangle graph from the vertex the
    previous support call ended on (mesh.py:12-139).  The fixture holds the outputs of the
    real reference with a fresh object per call; everything is bit-exact (contract: 1e-9)."""
    from distance3d_b200 import epa as d3epa, mpr as d3mpr
    cs, g = load_golden("meshgraph.npz")
    res = gjk.gjk_distance_batch(cs, g["pairs"]).cpu()
    ok = g["status"] <= 1
    assert np.max(np.abs(res["dist"][ok] - g["dist"][ok])) < TOL        # the stated tolerance
    assert np.max(np.abs(res["closest_a"][ok] - g["a"][ok])) < TOL
    assert np.max(np.abs(res["closest_b"][ok] - g["b"][ok])) < TOL
    assert np.array_equal(res["status"], g["status"])
    assert np.array_equal(res["dist"][ok], g["dist"][ok])               # and in fact bit-exact
    assert np.array_equal(res["closest_a"][ok], g["a"][ok])
    assert np.array_equal(res["closest_b"][ok], g["b"][ok])
    assert np.array_equal(res["iters"], g["iters"])
    for k in np.where(ok)[0]:
        n = res["n_points"][k]
        assert np.array_equal(res["simplex"][k, :n], g["Y"][k, :n])
    hit, it, _ = gjk.gjk_intersection_batch(cs, g["pairs"], want_iters=True)
    assert np.array_equal(hit.cpu().numpy(), g["hit"])
    compare_distance(cs, g["pairs"], exact_types=EXACT)
    # EPA and MPR climb the same graph
    sel = np.where(g["epa_status"] >= 0)[0]
    e = d3epa.epa_batch(cs, g["pairs"][sel], g["Y"][sel]).cpu()
    asserted = g["epa_status"][sel] == 7
    assert np.array_equal(e["status"] == 7, asserted)
    assert np.array_equal(e["mtv"][~asserted], g["epa_mtv"][sel][~asserted])
    assert np.array_equal(e["success"][~asserted], g["epa_success"][sel][~asserted])
    assert np.array_equal(e["n_faces"][~asserted], g["epa_n_faces"][sel][~asserted])
    # the thread-per-pair EPA kernel hands MeshGraph pairs to the warp kernel (hill-climbing state)
    monkeypatch.setenv("D3D_EPA_KERNEL", "thread")
    e2 = d3epa.epa_batch(cs, g["pairs"][sel], g["Y"][sel]).cpu()
    monkeypatch.delenv("D3D_EPA_KERNEL")
    assert int(e2["deferred"][0]) > 0
    for key in ("status", "mtv", "success", "n_faces", "iters"):
        assert np.array_equal(e2[key], e[key]), key
    m = d3mpr.mpr_batch(cs, g["pairs"], penetration=True)
    mh = m["hit"].cpu().numpy()
    assert np.array_equal(mh, g["mpr_hit"])
    h = mh.astype(bool)
    assert np.array_equal(m["depth"].cpu().numpy()[h], g["mpr_depth"][h])
    assert np.array_equal(m["dir"].cpu().numpy()[h], g["mpr_dir"][h])
    assert np.array_equal(m["pos"].cpu().numpy()[h], g["mpr_pos"][h])
    mi = d3mpr.mpr_batch(cs, g["pairs"], penetration=False)
    assert np.array_equal(mi["hit"].cpu().numpy(), g["mpr_hit_intersection"])


def test_meshgraph_object_caches_its_vertex_across_scalar_calls():
    """colliders.MeshGraph keeps the vertex of its last support call like the reference
    object (mesh.py:85): a sequence of support_function calls on ONE object reproduces the
    reference's sequence of points and cached indices."""
    from distance3d_b200 import colliders as C, mesh as d3mesh
    cs, g = load_golden("meshgraph.npz")
    t0 = 0
    for m in range(12):
        nt = g["tri_len"][m]
        tri = g["triangles"][t0:t0 + nt]
        o, l = cs.vert_off[m], cs.vert_len[m]
        obj = C.MeshGraph(cs.pose[m].copy(), cs.verts[o:o + l].copy(), tri)
        sf = d3mesh.MeshHillClimbingSupportFunction(cs.pose[m].copy(), cs.verts[o:o + l].copy(), tri)
        for t in range(g["seq_dirs"].shape[1]):
            assert np.array_equal(obj.support_function(g["seq_dirs"][m, t]), g["seq_pts"][m, t])
            assert obj._first_idx == g["seq_idx"][m, t]
            idx, pt = sf(g["seq_dirs"][m, t])
            assert idx == g["seq_idx"][m, t] and np.array_equal(pt, g["seq_pts"][m, t])
        t0 += nt
    t0 = sum(g["tri_len"][:12])
    # gjk() on objects: first call = fresh objects = the fixture; the second call starts from
    # the cached vertices and still agrees to the contract tolerance
    k = int(np.where((cs.type[g["pairs"][:, 0]] == 6) & (cs.type[g["pairs"][:, 1]] == 6)
                     & (g["status"] == 0))[0][0])
    objs = []
    for i in g["pairs"][k]:
        t0 = int(np.sum(g["tri_len"][:i]))
        o, l = cs.vert_off[i], cs.vert_len[i]
        objs.append(C.MeshGraph(cs.pose[i].copy(), cs.verts[o:o + l].copy(),
                                g["triangles"][t0:t0 + g["tri_len"][i]]))
    d, a, b, _ = gjk.gjk(*objs)
    assert d == g["dist"][k] and np.array_equal(a, g["a"][k]) and np.array_equal(b, g["b"][k])
    assert objs[0]._first_idx is not None and objs[1]._first_idx is not None
    d2, a2, b2, _ = gjk.gjk(*objs)
    assert abs(d2 - d) < TOL and np.max(np.abs(a2 - a)) < 1e-6


def test_meshgraph_random_batches_thread_and_warp_kernels():
    """Random MeshGraph colliders (shared meshes, 8-200 vertices, so both the thread and the
    warp kernel climb) against the oracle, bit-exact; arg-max fallback without a graph."""
    rs = np.random.RandomState(21)
    cs = d3random.random_meshgraph_set(rs, 40, 400, 200, hull_vertices=(8, 200), center_scale=1.2)
    pairs = d3random.random_pairs(rs, len(cs), 20000)
    res, ref = compare_distance(cs, pairs, exact_types=EXACT)
    assert 0.1 < np.mean(ref["dist"] == 0.0) < 0.9
    hit, _, _ = gjk.gjk_intersection_batch(cs, pairs)
    assert np.array_equal(hit.cpu().numpy(), O.gjk_intersection(cs, pairs, n_threads=O.max_threads())["hit"])
    nog = P.ColliderSet(cs.type, cs.pose, cs.param, cs.vert_off, cs.vert_len, cs.verts)
    compare_distance(nog, pairs[:4000], exact_types=EXACT)


@pytest.mark.parametrize("wire,slots", [(True, 3), (False, 2)])
def test_streamed_host_pipeline_equals_batch_call(wire, slots):
    """Host buffers in, host buffers out; wire=True ships the compact type-specific records
    (d3d_unpack_colliders expands them on the device), wire=False the 4x4-pose layout."""
    from distance3d_b200 import stream as d3stream
    rs = np.random.RandomState(15)
    pipe = d3stream.GjkDistanceStream(6000, 3000, 60000, slots=slots)
    with pytest.raises(ValueError):
        d3stream.GjkDistanceStream(10, 10, 10).submit(d3stream.pin_batch(
            d3random.random_collider_set(rs, 50), d3random.random_pairs(rs, 50, 5)))
    batches = []
    for b in range(5):
        cs = d3random.random_collider_set(rs, 2000 + 500 * b, names=d3random.PRIMITIVES + ("mesh", "cone"),
                                          hull_vertices=(4, 90))
        pairs = d3random.random_pairs(rs, len(cs), 1000 + 300 * b)
        batches.append((cs, pairs, d3stream.pin_batch(cs, pairs, wire=wire)))
    if wire:   # 99 B per collider on the primitive mix instead of 164 B
        prim = d3random.random_collider_set(rs, 5000)
        wt, wo, w = prim.wire()
        assert (wt.nbytes + wo.nbytes + w.nbytes) / len(prim) < 0.68 * 164
    tickets = []
    results = []
    for k, (cs, pairs, host) in enumerate(batches):
        tickets.append(pipe.submit(host))
        if k >= 1:
            results.append({n: v.clone() for n, v in pipe.result(tickets[k - 1]).items()})
    results.append({n: v.clone() for n, v in pipe.result(tickets[-1]).items()})
    for (cs, pairs, _), got in zip(batches, results):
        ref = O.gjk_distance(cs, pairs)
        assert np.array_equal(got["dist"].numpy(), ref["dist"])
        assert np.array_equal(got["closest_a"].numpy(), ref["a"])
        assert np.array_equal(got["closest_b"].numpy(), ref["b"])
        assert np.array_equal(got["status"].numpy(), ref["status"])


def test_wire_records_of_every_collider_type_unpack_exactly():
    """d3d_unpack_colliders reproduces the structure-of-arrays set bit for bit (all ten types)."""
    import ctypes
    import torch
    from distance3d_b200 import _lib, colliders as C
    rs = np.random.RandomState(3)
    cs = d3random.random_collider_set(rs, 400, names=d3random.PRIMITIVES + ("mesh", "cone"), hull_vertices=(4, 30))
    extra = P.pack_colliders([C.Disk(rs.randn(3), 0.7, np.array([0.0, 0.6, 0.8])),
                              C.Ellipse(rs.randn(3), np.eye(3)[:2], np.array([0.4, 0.9])),
                              C.MeshGraph(np.eye(4), rs.randn(12, 3), np.array([[0, 1, 2], [1, 2, 3]]))])
    extra.graph_off = extra.graph = extra.mesh_start = None
    cs = P.concat_sets([cs, extra])
    wt, wo, w = cs.wire()
    dev = torch.device("cuda")
    n = len(cs)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    out = dict(type=torch.empty(n, dtype=torch.int32, device=dev), pose=torch.empty((n, 4, 4), dtype=torch.float64, device=dev),
               param=torch.empty((n, 3), dtype=torch.float64, device=dev), vert_off=torch.empty(n, dtype=torch.int32, device=dev),
               vert_len=torch.empty(n, dtype=torch.int32, device=dev))
    a, b, c = t(wt), t(wo), t(w)
    _lib._check(_lib.lib().d3d_unpack_colliders(_lib.ptr(a), _lib.ptr(b), _lib.ptr(c), ctypes.c_int64(n),
                                                _lib.ptr(out["type"]), _lib.ptr(out["pose"]), _lib.ptr(out["param"]),
                                                _lib.ptr(out["vert_off"]), _lib.ptr(out["vert_len"]), _lib.stream_ptr()))
    assert np.array_equal(out["type"].cpu().numpy(), cs.type)
    assert np.array_equal(out["pose"].cpu().numpy(), cs.pose)
    used = cs.param.copy()
    assert np.array_equal(out["param"].cpu().numpy(), used)
    hv = np.isin(cs.type, (P.BOX, P.HULL, P.MESH))
    assert np.array_equal(out["vert_off"].cpu().numpy()[hv], cs.vert_off[hv])
    assert np.array_equal(out["vert_len"].cpu().numpy()[hv], cs.vert_len[hv])
    # the record offsets recomputed on the device from the type bytes (what the stream does instead of
    # uploading them): all ten types, set sizes around the 1024-collider tiles of the scan
    reps = cs
    for n_rep in (1, 3, 5300):
        big = P.concat_sets([cs] * n_rep) if n_rep > 1 else reps
        for cut in (len(big), 1023, 1024, 1025):
            if cut > len(big):
                continue
            sub_t = np.ascontiguousarray(big.type[:cut].astype(np.uint8))
            sizes = np.array([4, 14, 16, 15, 14, 1, 13, 7, 11, 14])[sub_t]
            expect = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
            d_t = t(sub_t)
            d_o = torch.full((cut,), -1, dtype=torch.int32, device=dev)
            _lib._check(_lib.lib().d3d_wire_offsets(_lib.ptr(d_t), ctypes.c_int64(cut), _lib.ptr(d_o), _lib.stream_ptr()))
            assert np.array_equal(d_o.cpu().numpy(), expect), (n_rep, cut)
    sizes0 = np.array([4, 14, 16, 15, 14, 1, 13, 7, 11, 14])[cs.type]
    assert np.array_equal(np.concatenate(([0], np.cumsum(sizes0)[:-1])).astype(np.int32), wo)


def test_fp32_mode_stated_tolerance():
    """Opt-in fp32 arithmetic: tolerance against the fp64 oracle as stated in include/d3d_b200.h:
    99.9 % of the distances within 1e-4 (unit-scale shapes), every valid result is an upper bound
    that is at most 0.2 off (rare early termination on hull-vs-curved pairs, as in Jolt's own
    single-precision GJK), intersection flags equal where the fp64 distance exceeds 1e-3."""
    rs = np.random.RandomState(16)
    cs = d3random.random_collider_set(rs, 4000, names=d3random.PRIMITIVES + ("mesh",))
    pairs = d3random.random_pairs(rs, len(cs), 60000)
    res = gjk.gjk_distance_batch(cs, pairs, dtype="f32").cpu()
    ref = O.gjk_distance(cs, pairs, n_threads=O.max_threads())
    ok = (ref["status"] <= 1) & (res["status"] <= 1)
    assert ok.mean() > 0.999
    err = np.abs(res["dist"][ok] - ref["dist"][ok])
    print("fp32 distance error: max %.3e p99.9 %.3e mean %.3e" % (err.max(), np.quantile(err, 0.999), err.mean()))
    assert np.quantile(err, 0.999) < 1e-4
    assert err.max() < 0.2
    assert np.mean(res["dist"][ok] - ref["dist"][ok] > -1e-3) == 1.0   # upper bound (up to contact flips)
    hit, _, _ = gjk.gjk_intersection_batch(cs, pairs, dtype="f32")
    ref_hit = O.gjk_intersection(cs, pairs, n_threads=O.max_threads())["hit"]
    hit = hit.cpu().numpy()
    far = ref["dist"] > 1e-3
    assert np.array_equal(hit[far], ref_hit[far])
    deep = ref["dist"] == 0.0
    assert (hit[deep] == ref_hit[deep]).mean() > 0.995   # grazing contacts may flip


def test_scalar_support_and_geometry_api_vs_reference_outputs():
    """geometry.support_function_*, containment.*_aabb, minkowski.support_function on single
    colliders reproduce the real reference's outputs stored in tests/golden/support.npz."""
    from distance3d_b200 import colliders as C, geometry, containment, minkowski, mesh
    cs, g = load_golden("support.npz")
    seen = set()
    for i in range(len(cs)):
        t = int(cs.type[i])
        if t in seen or (cs.margin is not None and cs.margin[i] != 0.0):
            continue
        seen.add(t)
        T, p, d = cs.pose[i], cs.param[i], g["dirs"][i, 0]
        if t == P.SPHERE:
            got = geometry.support_function_sphere(d, T[:3, 3], p[0])
            lo, hi = containment.sphere_aabb(T[:3, 3], p[0])
        elif t == P.CAPSULE:
            got = geometry.support_function_capsule(d, T, p[0], p[1])
            lo, hi = containment.capsule_aabb(T, p[0], p[1])
        elif t == P.CYLINDER:
            got = geometry.support_function_cylinder(d, T, p[0], p[1])
            lo, hi = containment.cylinder_aabb(T, p[0], p[1])
        elif t == P.ELLIPSOID:
            got = geometry.support_function_ellipsoid(d, T, p)
            lo, hi = containment.ellipsoid_aabb(T, p)
        elif t == P.CONE:
            got = geometry.support_function_cone(d, T, p[0], p[1])
            lo, hi = containment.cone_aabb(T, p[0], p[1])
        elif t == P.BOX:
            got = C.Box(T, p).support_function(d)
            lo, hi = containment.box_aabb(T, p)
            np.testing.assert_array_equal(geometry.convert_box_to_vertices(T, p).min(axis=0), lo)
        else:
            continue
        np.testing.assert_array_equal(got, g["support"][i, 0])
        np.testing.assert_array_equal(np.stack((lo, hi), axis=1), g["aabb"][i])
    assert len(seen) >= 6
    s1, s2 = C.Sphere(np.zeros(3), 1.0), C.Sphere(np.array([3.0, 0, 0]), 0.5)
    v, v1, v2 = minkowski.support_function(s1, s2, np.array([1.0, 0.0, 0.0]))
    np.testing.assert_allclose(v1, [1, 0, 0]); np.testing.assert_allclose(v2, [2.5, 0, 0])
    np.testing.assert_allclose(v, [-1.5, 0, 0])
    verts = np.random.RandomState(0).randn(30, 3)
    tri = mesh.make_convex_mesh(verts)
    idx, pt = mesh.MeshSupportFunction(np.eye(4), verts, tri)(np.array([0.0, 0.0, 1.0]))
    assert idx == int(np.argmax(verts[:, 2])) and np.allclose(pt, verts[idx])


def test_libccd_boolean_gjk_vs_reference_outputs_and_as_cross_check():
    """gjk_intersection_libccd (gjk/_gjk_libccd.py:14-266) batched: identical booleans to the
    reference's on the fixture (all collider types, Margin, MeshGraph), identical to the oracle
    on a large random batch, and - the role it has in the reference's tests
    (test_gjk.py:341-354) - an independent cross-check of the Jolt kernel."""
    for tag in ("far", "near"):
        cs, g = load_golden("libccd.npz", prefix=tag + "_cs_")
        hit, iters = gjk.gjk_intersection_libccd_batch(cs, g[tag + "_pairs"], want_iters=True)
        assert np.array_equal(hit.cpu().numpy(), g[tag + "_hit"])
        ref = O.gjk_intersection_libccd(cs, g[tag + "_pairs"])
        assert np.array_equal(iters.cpu().numpy(), ref["iters"])
    rs = np.random.RandomState(33)
    cs = d3random.random_collider_set(rs, 4000, names=d3random.PRIMITIVES + ("mesh", "cone"), center_scale=0.6,
                                      hull_vertices=(4, 60))
    pairs = d3random.random_pairs(rs, len(cs), 200000)
    hit, _ = gjk.gjk_intersection_libccd_batch(cs, pairs)
    hit = hit.cpu().numpy()
    ref = O.gjk_intersection_libccd(cs, pairs, n_threads=O.max_threads())
    assert np.array_equal(hit, ref["hit"])
    jolt, _, _ = gjk.gjk_intersection_batch(cs, pairs)
    dist = gjk.gjk_distance_batch(cs, pairs, want_points=False, want_simplex=False).dist.cpu().numpy()
    clear = (dist > 1e-6) | (dist == 0.0)
    disagree = hit != jolt.cpu().numpy()
    assert disagree[clear].mean() < 2e-4, disagree[clear].mean()   # touching / grazing pairs only
    assert 0.2 < hit.mean() < 0.8
    from distance3d_b200 import colliders as C
    s1, s2 = C.Sphere(np.zeros(3), 1.0), C.Sphere(np.array([0.0, 0.0, 1.5]), 1.0)
    assert gjk.gjk_intersection_libccd(s1, s2) is True
    assert gjk.gjk_intersection_libccd(s1, C.Sphere(np.array([0.0, 0.0, 2.5]), 1.0)) is False


@pytest.mark.parametrize("split_min", ["1", "100000000"])
def test_all_thread_instances_and_the_warp_kernel_in_one_batch(monkeypatch, split_min):
    """k_bin_scan sends a class of type bins (primitives / + hulls / everything) to its own
    instance of the thread kernel only when the class is large; D3D_GJK_SPLIT_MIN=1 makes a small
    batch run on all three instances plus the warp kernel (hulls above 64 vertices), a huge value
    merges everything into one launch.  Both must reproduce the oracle bit for bit."""
    monkeypatch.setenv("D3D_GJK_SPLIT_MIN", split_min)
    rs = np.random.RandomState(77)
    cs = d3random.random_collider_set(rs, 900, names=d3random.PRIMITIVES + ("mesh", "cone"), center_scale=0.9,
                                      hull_vertices=(4, 120))
    pairs = d3random.random_pairs(rs, len(cs), 9000)
    res, ref = compare_distance(cs, pairs, exact_types=EXACT)
    hit, _, _ = gjk.gjk_intersection_batch(cs, pairs, want_iters=True)
    assert np.array_equal(hit.cpu().numpy().astype(bool), O.gjk_intersection(cs, pairs, n_threads=O.max_threads())["hit"].astype(bool))
