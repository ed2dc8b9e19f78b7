"""Shared helpers for the tests: golden fixtures and random packed colliders."""
import os

import numpy as np

from distance3d_b200.pack import ColliderSet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, prefix="cs_"):
    d = dict(np.load(os.path.join(GOLDEN, name)))
    p = prefix
    cs = ColliderSet(d[p + "type"], d[p + "pose"], d[p + "param"], d[p + "vert_off"], d[p + "vert_len"],
                     d[p + "verts"], d.get(p + "margin"), graph_off=d.get(p + "graph_off"),
                     graph=d.get(p + "graph"), mesh_start=d.get(p + "mesh_start"))
    return cs, d


def polygon_boundary_distance(a, b):
    """Largest distance of a vertex of polygon `a` from the boundary (edges) of polygon `b`."""
    worst = 0.0
    nb = len(b)
    for p in a:
        best = np.inf
        for k in range(nb):
            s, e = b[k], b[(k + 1) % nb]
            d = e - s
            L = float(np.dot(d, d))
            t = 0.0 if L == 0.0 else min(1.0, max(0.0, float(np.dot(p - s, d)) / L))
            best = min(best, float(np.linalg.norm(p - (s + t * d))))
        worst = max(worst, best)
    return worst


def polygon_area_centroid(v):
    """Area and centroid of a planar polygon in 3-D (fan triangulation)."""
    area, com = 0.0, np.zeros(3)
    for k in range(1, len(v) - 1):
        a = 0.5 * np.linalg.norm(np.cross(v[k] - v[0], v[k + 1] - v[0]))
        area += a
        com += a * (v[0] + v[k] + v[k + 1]) / 3.0
    return area, (com / area if area > 0 else v.mean(axis=0))


def compare_tetra_results(res, g, key):
    """Geometric parity of intersect_tetrahedron_pairs outputs with the reference's: same
    booleans, contact planes within 1e-9, contact polygons equal as REGIONS within 1e-9 (vertex
    lists may differ by redundant collinear vertices: where two coplanar faces meet, the
    reference intersects two numerically parallel half-planes at a noise-determined point on
    their common line).  Returns the number of regular intersecting pairs compared."""
    hit = g[key + "hit"]
    assert np.array_equal(res["hit"], hit)
    m = hit == 1
    assert np.max(np.abs(res["plane"][m] - g[key + "plane"][m]), initial=0.0) < 1e-9
    n = 0
    for q in np.where(m & (res["status"] == 0))[0]:
        a = res["polygon"][q, :res["n_vertices"][q]]
        b = g[key + "poly"][q, :g[key + "nv"][q]]
        assert polygon_boundary_distance(a, b) < 1e-9 and polygon_boundary_distance(b, a) < 1e-9
        (aa, ca), (ab, cb) = polygon_area_centroid(a), polygon_area_centroid(b)
        assert abs(aa - ab) < 1e-9 and np.max(np.abs(ca - cb)) < 1e-7
        n += 1
    same = m & (res["status"] == 1)   # "same tetrahedron" branch: three copies of one point
    assert np.max(np.abs(res["polygon"][same, :3] - g[key + "poly"][same, :3]), initial=0.0) < 1e-9
    return n
