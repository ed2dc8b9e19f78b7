"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the REAL reference.

Runs only in the build container (imports /root/reference through
oracle/refshim).  The fixtures store the packed inputs (ColliderSet arrays)
and the outputs of the reference's own functions:

  gjk.npz      gjk.gjk / gjk.gjk_intersection / gjk_distance_jolt_iterations
  epa.npz      epa.epa on the GJK simplices of intersecting pairs
  mpr.npz      mpr.mpr_penetration / mpr_intersection
  aabb.npz     collider.aabb(), AabbTree.overlaps_aabb_tree, all_aabbs_overlap
  support.npz  collider.support_function on random directions
  hulls.npz    ConvexHullVertices with 64-256 vertices (config C3 shape)
  epa_degenerate.npz  epa.epa on simplices with duplicated / 1e-9-apart points (--only-epa-degenerate)

Usage: python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refbridge  # noqa: E402

refbridge.setup()
from distance3d import gjk, epa, mpr, aabb_tree  # noqa: E402
from distance3d.gjk._gjk_jolt import gjk_distance_jolt_iterations  # noqa: E402

OUT = os.path.join(refbridge.REPO, "tests", "golden")
MAX_FLOAT = np.finfo(float).max
ALL = ["sphere", "ellipsoid", "capsule", "cylinder", "box", "mesh", "cone", "disk", "ellipse"]


def set_arrays(cs, prefix="cs_"):
    d = {prefix + k: getattr(cs, k) for k in ("type", "pose", "param", "vert_off", "vert_len", "verts")}
    for k in ("margin", "graph_off", "graph", "mesh_start"):
        if getattr(cs, k) is not None:
            d[prefix + k] = getattr(cs, k)
    return d


def run_gjk(cols, pairs):
    n = len(pairs)
    dist = np.zeros(n); a = np.zeros((n, 3)); b = np.zeros((n, 3)); Y = np.zeros((n, 4, 3))
    status = np.zeros(n, dtype=np.int32); iters = np.zeros(n, dtype=np.int32)
    hit = np.zeros(n, dtype=np.uint8)
    for k, (i, j) in enumerate(pairs):
        try:
            d, pa, pb, yy = gjk.gjk(cols[i], cols[j])
        except AssertionError:
            status[k] = 4
            continue
        iters[k] = gjk_distance_jolt_iterations(cols[i], cols[j])
        if pa is None:
            dist[k] = MAX_FLOAT; status[k] = 3
        else:
            dist[k] = d; a[k] = pa; b[k] = pb; Y[k] = yy
            status[k] = 1 if d == 0.0 else 0
        hit[k] = gjk.gjk_intersection(cols[i], cols[j])
    return dict(dist=dist, a=a, b=b, Y=Y, status=status, iters=iters, hit=hit)


def fresh(c):
    """New reference object for the same shape: a MeshGraph carries hidden state (the
    vertex its last support call ended on, mesh.py:85), so every reference call of the
    MeshGraph fixture gets objects in their initial state and the fixture does not
    depend on the order of the calls."""
    from distance3d import colliders as RC
    if isinstance(c, RC.MeshGraph):
        return RC.MeshGraph(c.mesh2origin.copy(), c.vertices, c.triangles)
    return c


def meshgraph_fixture():
    """MeshGraph: hill climbing with vertex caching (mesh.py:12-139)."""
    from distance3d import colliders as RC
    from distance3d_b200.mesh import build_mesh_graph
    rsm = np.random.RandomState(99)
    meshes = []
    for k in range(48):  # 8 .. 200 vertices; close together so that EPA / MPR get hits
        nv = int(rsm.choice([8, 12, 20, 30, 60, 120, 200]))
        meshes += refbridge.random_reference_colliders(
            rsm, 1, ["mesh"], hull_as_vertices=False, mesh=dict(n_vertices=nv, center_scale=[0.5, 1.5][k % 2]))
    names = ["sphere", "capsule", "box", "cylinder", "ellipsoid"]
    others = refbridge.random_reference_colliders(
        rsm, 32, names, **{n: dict(center_scale=1.2) for n in names})
    cols = meshes + others
    # the adjacency record built by the product equals the reference object's tables
    for c in meshes:
        g = build_mesh_graph(c.vertices, c.triangles)
        sf = c._support_function
        assert g[0] == sf.first_idx and list(g[1:7]) == list(sf.shortcut_connections)
        for v in range(len(c.vertices)):
            mine = list(g[g[7 + v]:g[8 + v]])
            assert mine == (list(sf.connections[v]) if v in sf.connections else []), v
    cs = refbridge.to_set(cols)
    n_pairs = 400
    pairs = np.stack([rsm.randint(0, 48, n_pairs), rsm.randint(0, 80, n_pairs)], axis=1).astype(np.int32)
    swap = rsm.rand(n_pairs) < 0.3  # the mesh is not always collider 1
    pairs[swap] = pairs[swap][:, ::-1]
    n = n_pairs
    dist = np.zeros(n); a = np.zeros((n, 3)); b = np.zeros((n, 3)); Y = np.zeros((n, 4, 3))
    status = np.zeros(n, dtype=np.int32); iters = np.zeros(n, dtype=np.int32)
    hit = np.zeros(n, dtype=np.uint8)
    mtv = np.zeros((n, 3)); esuccess = np.zeros(n, dtype=np.uint8); nfaces = np.zeros(n, dtype=np.int32)
    estatus = np.full(n, -1, dtype=np.int32)
    mhit = np.zeros(n, dtype=np.uint8); mhit_i = np.zeros(n, dtype=np.uint8)
    depth = np.zeros(n); pdir = np.zeros((n, 3)); pos = np.zeros((n, 3))
    for k, (i, j) in enumerate(pairs):
        d, pa, pb, yy = gjk.gjk(fresh(cols[i]), fresh(cols[j]))
        iters[k] = gjk_distance_jolt_iterations(fresh(cols[i]), fresh(cols[j]))
        if pa is None:
            dist[k] = MAX_FLOAT; status[k] = 3
        else:
            dist[k] = d; a[k] = pa; b[k] = pb; Y[k] = yy
            status[k] = 1 if d == 0.0 else 0
        hit[k] = gjk.gjk_intersection(fresh(cols[i]), fresh(cols[j]))
        if status[k] == 1:
            try:
                m, faces, ok = epa.epa(Y[k].copy(), fresh(cols[i]), fresh(cols[j]))
                estatus[k] = 1; mtv[k] = m; esuccess[k] = ok; nfaces[k] = len(faces)
            except AssertionError:
                estatus[k] = 7
        h, dpt, dr, ps = mpr.mpr_penetration(fresh(cols[i]), fresh(cols[j]))
        mhit[k] = h
        mhit_i[k] = mpr.mpr_intersection(fresh(cols[i]), fresh(cols[j]))
        if h:
            depth[k] = dpt; pdir[k] = dr; pos[k] = ps
    # one object, a sequence of support calls: the start vertex is carried from call to call
    seq_dirs = rsm.randn(48, 12, 3)
    seq_dirs[:, 5] = seq_dirs[:, 4] * (1.0 + 1e-15)  # nearly repeated direction
    seq_pts = np.zeros((48, 12, 3)); seq_idx = np.zeros((48, 12), dtype=np.int32)
    for m_i, c in enumerate(meshes):
        obj = fresh(c)
        for t in range(12):
            seq_pts[m_i, t] = obj.support_function(np.ascontiguousarray(seq_dirs[m_i, t]))
            seq_idx[m_i, t] = obj._support_function.first_idx
    tri_len = np.array([len(c.triangles) for c in meshes], dtype=np.int32)
    print("meshgraph: %d intersecting, %d epa ok, %d epa max_faces, %d mpr hits, iters mean %.1f" % (
        (status == 1).sum(), (estatus == 1).sum(), (estatus == 7).sum(), mhit.sum(), iters.mean()))
    np.savez_compressed(
        os.path.join(OUT, "meshgraph.npz"), pairs=pairs, dist=dist, a=a, b=b, Y=Y, status=status,
        iters=iters, hit=hit, epa_mtv=mtv, epa_success=esuccess, epa_n_faces=nfaces,
        epa_status=estatus, mpr_hit=mhit, mpr_hit_intersection=mhit_i, mpr_depth=depth,
        mpr_dir=pdir, mpr_pos=pos, seq_dirs=seq_dirs, seq_pts=seq_pts, seq_idx=seq_idx,
        triangles=np.concatenate([np.asarray(c.triangles) for c in meshes]).astype(np.int32),
        tri_len=tri_len, **set_arrays(cs))


def branched_fixture():
    """Self-collision of a branched robot with the reference's detect(): white-lists are
    not symmetric there (urdf_utils.py:79-81 keeps one child per link)."""
    from pytransform3d.urdf import UrdfTransformManager
    import distance3d.broad_phase
    from distance3d import self_collision
    urdf_path = os.path.join(refbridge.REPO, "tests", "data", "robot_branched.urdf")
    tm = UrdfTransformManager()
    with open(urdf_path) as f:
        tm.load_urdf(f.read(), mesh_path=os.path.dirname(urdf_path))
    bvh = distance3d.broad_phase.BoundingVolumeHierarchy(tm, "robot_branched")
    bvh.fill_tree_with_colliders(tm, make_artists=False, fill_self_collision_whitelists=True)
    joints = ["j_shoulder_l", "j_arm_l", "j_shoulder_r", "j_arm_r", "j_head"]
    rsq = np.random.RandomState(17)
    q = rsq.uniform(-np.pi, np.pi, size=(300, len(joints)))
    q[0] = 0.0
    frames = list(bvh.colliders_.keys())
    wl = bvh.self_collision_whitelists_
    masks = np.zeros((len(q), len(frames)), dtype=np.uint8)
    either = np.zeros_like(masks)   # pair tested unless BOTH sides white-list it, a hit marks both
    anyhit = np.zeros(len(q), dtype=np.uint8)
    for b in range(len(q)):
        for j, name in enumerate(joints):
            tm.set_joint(name, q[b, j])
        bvh.update_collider_poses()
        contacts = self_collision.detect(bvh)
        masks[b] = [contacts[fr] for fr in frames]
        anyhit[b] = self_collision.detect_any(bvh)
        for i, fi in enumerate(frames):
            for j2 in range(i + 1, len(frames)):
                fj = frames[j2]
                if fj in wl[fi] and fi in wl[fj]:
                    continue
                ci, cj = bvh.colliders_[fi], bvh.colliders_[fj]
                a1, a2 = ci.aabb(), cj.aabb()
                if np.all((a1[:, 0] <= a2[:, 1]) & (a1[:, 1] >= a2[:, 0])) and gjk.gjk_intersection(ci, cj):
                    either[b, i] = either[b, j2] = 1
    asym = [(fi, fj) for fi in frames for fj in frames if fj in wl[fi] and fi not in wl[fj]]
    print("branched: %d frames, %d asymmetric white-list entries, colliding configs %d/%d, "
          "detect() vs pairwise rule: %d differing mask bits" % (
              len(frames), len(asym), (masks.sum(1) > 0).sum(), len(q), (masks != either).sum()))
    np.savez_compressed(os.path.join(OUT, "self_collision_branched.npz"), q=q, mask=masks,
                        any=anyhit, frames=np.array(frames), joints=np.array(joints),
                        wl_keys=np.array(frames),
                        wl=np.array([[int(fj in wl[fi]) for fj in frames] for fi in frames], dtype=np.uint8))


def _load_hydroelastic():
    """The reference's hydroelastic modules without its package __init__ (which imports the
    matplotlib / Open3D visualisation): only the numerical files are loaded."""
    import importlib.util
    import types
    pkg_dir = os.path.join(refbridge.REFERENCE, "distance3d", "hydroelastic_contact")
    pkg = types.ModuleType("distance3d.hydroelastic_contact")
    pkg.__path__ = [pkg_dir]
    sys.modules["distance3d.hydroelastic_contact"] = pkg
    mods = {}
    for name in ("_halfplanes", "_barycentric_transform", "_mesh_processing", "_tetra_mesh_creation",
                 "_tetrahedron_intersection"):
        spec = importlib.util.spec_from_file_location("distance3d.hydroelastic_contact." + name,
                                                      os.path.join(pkg_dir, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def tetra_fixture():
    """Hydroelastic broad-phase consumer: two tetrahedral meshes made by the
    reference's own mesh generators, broad phase with all_aabbs_overlap, narrow phase with
    intersect_tetrahedron_pairs (hydroelastic_contact/_interface.py:52-101 without the forces)."""
    H = _load_hydroelastic()
    from distance3d.utils import transform_points
    from pytransform3d import transformations as pt
    mk = H["_tetra_mesh_creation"]
    rs = np.random.RandomState(123)
    cases = []
    bodies = [("sphere", mk.make_tetrahedral_sphere(0.15, 2)), ("box", mk.make_tetrahedral_box(np.array([0.2, 0.25, 0.3]))),
              ("cube", mk.make_tetrahedral_cube(0.2)), ("ellipsoid", mk.make_tetrahedral_ellipsoid(np.array([0.1, 0.15, 0.2]), 2)),
              ("cylinder", mk.make_tetrahedral_cylinder(0.1, 0.3, 0.05))]
    out = {}
    n_case = 0
    for a in range(len(bodies)):
        for b in range(a, len(bodies)):
            (na, (va, ta, pa)), (nb, (vb, tb, pb)) = bodies[a], bodies[b]
            A2o = pt.random_transform(rs); A2o[:3, 3] *= 0.03
            B2o = pt.random_transform(rs); B2o[:3, 3] *= 0.03
            # rigid body 1 expressed in the frame of body 2 (_interface.py:74)
            rel = np.dot(np.linalg.inv(B2o), A2o)
            if a == b and n_case % 2 == 0:
                rel = np.eye(4)       # bit-identical bodies: the "same tetrahedron" branch (:132-134)
            tp1 = np.ascontiguousarray(transform_points(rel, va)[ta])
            tp2 = np.ascontiguousarray(vb[tb])
            e1 = np.ascontiguousarray(pa[ta]); e2 = np.ascontiguousarray(pb[tb])
            aabbs1 = H["_mesh_processing"].tetrahedral_mesh_aabbs(tp1)
            aabbs2 = H["_mesh_processing"].tetrahedral_mesh_aabbs(tp2)
            b1, b2, bpairs = aabb_tree.all_aabbs_overlap(aabbs1, aabbs2)
            X1 = H["_barycentric_transform"].barycentric_transforms(tp1[b1])
            X2 = H["_barycentric_transform"].barycentric_transforms(tp2[b2])
            X1d = {j: X1[i] for i, j in enumerate(b1)}
            X2d = {j: X2[i] for i, j in enumerate(b2)}
            ym1, ym2 = 1.0 + (n_case % 3), 1.0
            inter, planes, polys, i1, i2 = H["_tetrahedron_intersection"].intersect_tetrahedron_pairs(
                bpairs, tp1, tp2, e1, e2, X1d, X2d, ym1, ym2)
            bp = np.array(bpairs, dtype=np.int32).reshape(-1, 2)
            hit = np.zeros(len(bp), dtype=np.uint8)
            plane = np.zeros((len(bp), 4)); nv = np.zeros(len(bp), dtype=np.int32); poly = np.zeros((len(bp), 12, 3))
            lookup = {tuple(p): k for k, p in enumerate(bp.tolist())}
            for q, (i, j) in enumerate(zip(i1, i2)):
                k = lookup[(int(i), int(j))]
                hit[k] = 1; plane[k] = planes[q]; nv[k] = len(polys[q]); poly[k, :len(polys[q])] = polys[q][:12]
            key = "c%d_" % n_case
            out.update({key + "tp1": tp1, key + "tp2": tp2, key + "e1": e1, key + "e2": e2,
                        key + "aabb1": aabbs1, key + "aabb2": aabbs2, key + "pairs": bp, key + "hit": hit,
                        key + "plane": plane, key + "nv": nv, key + "poly": poly,
                        key + "ym": np.array([ym1, ym2])})
            print("tetra case %d %s-%s: %d x %d tetrahedra, %d broad pairs, %d intersecting, max polygon %d" % (
                n_case, na, nb, len(tp1), len(tp2), len(bp), int(hit.sum()), int(nv.max()) if len(nv) else 0))
            n_case += 1
    out["n_cases"] = np.array(n_case)
    np.savez_compressed(os.path.join(OUT, "tetra.npz"), **out)


def libccd_fixture():
    """gjk_intersection_libccd (gjk/_gjk_libccd.py:14-266) on all collider types + Margin +
    MeshGraph, far apart and close together; a fresh reference object per call."""
    from distance3d import colliders as RC
    from distance3d.gjk import gjk_intersection_libccd
    out = {}
    for tag, scale in (("far", 1.0), ("near", 0.3)):
        rs = np.random.RandomState(77)
        kw = {n: dict(center_scale=scale) for n in ALL if n not in ("disk", "ellipse")}
        cols = refbridge.random_reference_colliders(rs, 300, ALL, **kw)
        cols += refbridge.random_reference_colliders(rs, 60, ["mesh"], hull_as_vertices=False,
                                                     mesh=dict(center_scale=scale, n_vertices=24))
        for k in range(0, 300, 15):
            cols[k] = RC.Margin(cols[k], 0.1 * rs.rand())
        cs = refbridge.to_set(cols)
        pairs = rs.randint(0, len(cols), size=(2500, 2)).astype(np.int32)
        hit = np.array([gjk_intersection_libccd(fresh(cols[i]), fresh(cols[j])) for i, j in pairs], dtype=np.uint8)
        jolt = np.array([gjk.gjk_intersection(fresh(cols[i]), fresh(cols[j])) for i, j in pairs], dtype=np.uint8)
        print("libccd %s: %d pairs, %.3f intersecting, %d disagree with the Jolt variant" % (
            tag, len(pairs), hit.mean(), int((hit != jolt).sum())))
        out.update({tag + "_pairs": pairs, tag + "_hit": hit, tag + "_hit_jolt": jolt})
        out.update(set_arrays(cs, prefix=tag + "_cs_"))
    np.savez_compressed(os.path.join(OUT, "libccd.npz"), **out)


def epa_degenerate_fixture():
    """epa_degenerate.npz: the real reference's epa.epa on GJK simplices with a duplicated point or two
    points 1e-9 apart (closer than the edge-matching epsilon of epa.py:189-191): the region where
    matching loose edges by coordinates differs from matching them by identity."""
    rs = np.random.RandomState(909)
    names = ["sphere", "ellipsoid", "capsule", "cylinder", "box", "mesh", "cone"]
    cols = refbridge.random_reference_colliders(rs, 200, names, **{n: dict(center_scale=0.3) for n in names})
    cs = refbridge.to_set(cols)
    pairs = rs.randint(0, len(cols), size=(900, 2)).astype(np.int32)
    g = run_gjk(cols, pairs)
    sel = np.where((g["dist"] == 0.0) & (g["status"] == 1))[0]
    pairs, Y = pairs[sel], g["Y"][sel].copy()
    kind = np.zeros(len(sel), dtype=np.int32)          # 0 untouched, 1 duplicated point, 2 near-duplicate
    Y[::3, 1] = Y[::3, 0]; kind[::3] = 1
    Y[1::3, 2] = Y[1::3, 3] + 1e-9; kind[1::3] = 2
    n = len(sel)
    mtv = np.zeros((n, 3)); success = np.zeros(n, dtype=np.uint8); nfaces = np.zeros(n, dtype=np.int32)
    status = np.ones(n, dtype=np.int32)
    faces_all = np.zeros((n, 64, 4, 3))
    for k, (i, j) in enumerate(pairs):
        try:
            with np.errstate(all="ignore"):
                m, faces, ok = epa.epa(Y[k].copy(), cols[i], cols[j])
        except AssertionError:
            status[k] = 7
            continue
        mtv[k] = m; success[k] = ok; nfaces[k] = len(faces); faces_all[k, :len(faces)] = faces
    print("epa_degenerate: %d simplices, %d max_faces assertions, %d converged" % (n, int((status == 7).sum()), int(success.sum())))
    np.savez_compressed(os.path.join(OUT, "epa_degenerate.npz"), pairs=pairs, Y=Y, kind=kind, mtv=mtv,
                        success=success, n_faces=nfaces, status=status, faces=faces_all, **set_arrays(cs))


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--only-epa-degenerate" in sys.argv:
        epa_degenerate_fixture()
        return
    if "--only-libccd" in sys.argv:
        libccd_fixture()
        return
    if "--only-tetra" in sys.argv:
        tetra_fixture()
        return
    if "--only-meshgraph" in sys.argv:
        meshgraph_fixture()
        return
    if "--only-branched" in sys.argv:
        branched_fixture()
        return
    rs = np.random.RandomState(2024)

    # ---- GJK on all 9 collider types (+ Margin) ---------------------------
    from distance3d import colliders as RC
    cols = refbridge.random_reference_colliders(rs, 240, ALL)
    for k in range(0, 240, 12):
        cols[k] = RC.Margin(cols[k], 0.1 * rs.rand())
    cs = refbridge.to_set(cols)
    pairs = rs.randint(0, len(cols), size=(600, 2)).astype(np.int32)
    res = run_gjk(cols, pairs)
    # n_points is not returned by the reference; the oracle provides it and the
    # test only compares the first n_points rows of Y
    np.savez_compressed(os.path.join(OUT, "gjk.npz"), pairs=pairs, **set_arrays(cs), **res)

    # support + aabb + center on the same set
    dirs = rs.randn(len(cols), 4, 3)
    dirs[:, 3] = [1.0, 0.0, 0.0]
    sup = np.array([[c.support_function(np.ascontiguousarray(d)) for d in dirs[i]]
                    for i, c in enumerate(cols)])
    aabbs = np.array([c.aabb() for c in cols])
    centers = np.array([c.center() for c in cols])
    np.savez_compressed(os.path.join(OUT, "support.npz"), dirs=dirs, support=sup, aabb=aabbs,
                        center=centers, **set_arrays(cs))

    # ---- close-together shapes: EPA + MPR ---------------------------------
    names = ["sphere", "ellipsoid", "capsule", "cylinder", "box", "mesh", "cone"]
    cols2 = refbridge.random_reference_colliders(
        rs, 160, names, **{n: dict(center_scale=0.3) for n in names})
    cs2 = refbridge.to_set(cols2)
    pairs2 = rs.randint(0, len(cols2), size=(400, 2)).astype(np.int32)
    g2 = run_gjk(cols2, pairs2)
    n = len(pairs2)
    mtv = np.zeros((n, 3)); success = np.zeros(n, dtype=np.uint8); nfaces = np.zeros(n, dtype=np.int32)
    estatus = np.full(n, -1, dtype=np.int32)
    faces_all = np.zeros((n, 64, 4, 3))
    for k, (i, j) in enumerate(pairs2):
        if g2["dist"][k] != 0.0 or g2["status"][k] != 1:
            continue
        try:
            m, faces, ok = epa.epa(g2["Y"][k].copy(), cols2[i], cols2[j])
        except AssertionError:
            estatus[k] = 7
            continue
        estatus[k] = 1
        mtv[k] = m; success[k] = ok; nfaces[k] = len(faces); faces_all[k, :len(faces)] = faces
    np.savez_compressed(os.path.join(OUT, "epa.npz"), pairs=pairs2, Y=g2["Y"], gjk_dist=g2["dist"],
                        mtv=mtv, success=success, n_faces=nfaces, status=estatus,
                        faces=faces_all.astype(np.float64), **set_arrays(cs2))
    hit = np.zeros(n, dtype=np.uint8); hit_i = np.zeros(n, dtype=np.uint8)
    depth = np.zeros(n); pdir = np.zeros((n, 3)); pos = np.zeros((n, 3))
    for k, (i, j) in enumerate(pairs2):
        h, dpt, dr, ps = mpr.mpr_penetration(cols2[i], cols2[j])
        hit[k] = h
        hit_i[k] = mpr.mpr_intersection(cols2[i], cols2[j])
        if h:
            depth[k] = dpt; pdir[k] = dr; pos[k] = ps
    np.savez_compressed(os.path.join(OUT, "mpr.npz"), pairs=pairs2, hit=hit, hit_intersection=hit_i,
                        depth=depth, dir=pdir, pos=pos, gjk_hit=g2["hit"], **set_arrays(cs2))

    # ---- hulls with 64-256 vertices (config C3 shape) ----------------------
    cols3 = []
    for _ in range(24):
        nv = rs.randint(64, 257)
        cols3.extend(refbridge.random_reference_colliders(
            rs, 1, ["mesh"], mesh=dict(n_vertices=nv, center_scale=0.8)))
    cs3 = refbridge.to_set(cols3)
    pairs3 = np.array([(i, j) for i in range(24) for j in range(24) if i < j][:120], dtype=np.int32)
    g3 = run_gjk(cols3, pairs3)
    n = len(pairs3)
    mtv = np.zeros((n, 3)); success = np.zeros(n, dtype=np.uint8); nfaces = np.zeros(n, dtype=np.int32)
    estatus = np.full(n, -1, dtype=np.int32)
    for k, (i, j) in enumerate(pairs3):
        if g3["dist"][k] != 0.0:
            continue
        try:
            m, faces, ok = epa.epa(g3["Y"][k].copy(), cols3[i], cols3[j])
        except AssertionError:
            estatus[k] = 7
            continue
        estatus[k] = 1; mtv[k] = m; success[k] = ok; nfaces[k] = len(faces)
    np.savez_compressed(os.path.join(OUT, "hulls.npz"), pairs=pairs3, epa_mtv=mtv, epa_success=success,
                        epa_n_faces=nfaces, epa_status=estatus, **set_arrays(cs3), **g3)

    # ---- broad phase -------------------------------------------------------
    from distance3d.random import rand_capsule
    rs32 = np.random.RandomState(32)
    caps = []
    for _ in range(1500):
        caps.append(RC.Capsule(*rand_capsule(rs32, center_scale=2.0, radius_scale=0.1, height_scale=0.5)))
    cs4 = refbridge.to_set(caps)
    A = np.array([c.aabb() for c in caps])
    tree = aabb_tree.AabbTree()
    tree.insert_aabbs(A)
    _, _, _, tpairs = tree.overlaps_aabb_tree(tree)
    tpairs = np.array(tpairs, dtype=np.int32).reshape(-1, 2)
    _, _, bpairs = aabb_tree.all_aabbs_overlap(A[:300], A[300:700])
    bpairs = np.array(bpairs, dtype=np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "aabb.npz"), aabb=A, tree_pairs=tpairs, brute_pairs=bpairs,
                        tree_nodes=tree.nodes, tree_root=tree.root, **set_arrays(cs4))
    # ---- robot self-collision (config C4 shape): reference detect() over random joints
    from pytransform3d.urdf import UrdfTransformManager
    import distance3d.broad_phase
    from distance3d import self_collision
    urdf_path = os.path.join(refbridge.REPO, "tests", "data", "robot_arm.urdf")
    tm = UrdfTransformManager()
    with open(urdf_path) as f:
        tm.load_urdf(f.read(), mesh_path=os.path.dirname(urdf_path))
    bvh = distance3d.broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
    bvh.fill_tree_with_colliders(tm, make_artists=False, fill_self_collision_whitelists=True)
    rsq = np.random.RandomState(7)
    q = rsq.uniform(-np.pi, np.pi, size=(150, 6))
    q[0] = 0.0
    q[1] = [0.0, 1.57, 1.57, 0.0, 1.93, 0.0]   # distance3d/test/test_self_collision.py:26-33 -> 0 contacts
    q[2] = [0.0, 1.57, 1.57, 0.0, 2.05, 0.0]   # :35-42 -> 3 contacts
    frames = list(bvh.colliders_.keys())
    masks = np.zeros((len(q), len(frames)), dtype=np.uint8)
    poses = np.zeros((len(q), len(frames), 4, 4))
    for b in range(len(q)):
        for j in range(6):
            tm.set_joint("joint%d" % (j + 1), q[b, j])
        bvh.update_collider_poses()
        contacts = self_collision.detect(bvh)
        masks[b] = [contacts[fr] for fr in frames]
        poses[b] = [tm.get_transform(fr, "origin") for fr in frames]
    assert masks[0].sum() == 0 and masks[1].sum() == 0 and masks[2].sum() == 3
    np.savez_compressed(os.path.join(OUT, "self_collision.npz"), q=q, mask=masks, poses=poses,
                        frames=np.array(frames))

    meshgraph_fixture()
    branched_fixture()
    tetra_fixture()
    libccd_fixture()
    libccd_fixture()

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
