"""CPU: known-answer tests restated from the reference's own test-suite
(numbers copied from distance3d/test/*.py), run against the oracle."""
import numpy as np
import pytest

from distance3d_b200 import colliders as C
from distance3d_b200.pack import pack_colliders
from distance3d_b200._transforms import (
    transform_from, active_matrix_from_extrinsic_euler_xyz)
from oracle import cpu_oracle as O

MAX_FLOAT = np.finfo(float).max


def dist(c1, c2, **kw):
    cs = pack_colliders([c1, c2])
    return O.gjk_distance(cs, [[0, 1]], **kw)


def test_gjk_boxes_known_answer():
    # distance3d/test/test_gjk.py:104-126
    box2origin = np.array([
        [-0.29265666, -0.76990535, 0.56709596, 0.1867558],
        [0.93923897, -0.12018753, 0.32153556, -0.09772779],
        [-0.17939408, 0.62673815, 0.75829879, 0.09500884],
        [0., 0., 0., 1.]])
    size = np.array([2.89098828, 1.15032456, 2.37517511])
    box2origin2 = np.array([
        [-0.29265666, -0.76990535, 0.56709596, 3.73511598],
        [0.93923897, -0.12018753, 0.32153556, -1.95455576],
        [-0.17939408, 0.62673815, 0.75829879, 1.90017684],
        [0., 0., 0., 1.]])
    size2 = np.array([0.96366276, 0.38344152, 0.79172504])
    r = dist(C.Box(box2origin, size), C.Box(box2origin2, size2))
    assert r["dist"][0] == pytest.approx(1.7900192730149391)
    hit = O.gjk_intersection(pack_colliders([C.Box(box2origin, size), C.Box(box2origin2, size2)]),
                             [[0, 1]])
    assert hit["hit"][0] == 0


EPA_VERTICES1 = np.array([
    [1.76405235, 0.40015721, 0.97873798],
    [2.2408932, 1.86755799, -0.97727788],
    [0.4105985, 0.14404357, 1.45427351],
    [0.33367433, 1.49407907, -0.20515826],
    [0.3130677, -0.85409574, -2.55298982],
    [2.26975462, -1.45436567, 0.04575852],
    [-0.18718385, 1.53277921, 1.46935877]])
EPA_VERTICES2 = np.array([
    [-2.32605299, -0.31242692, 0.07599278],
    [0.88503416, 1.23786508, -0.467683],
    [-0.64755927, -1.01306774, -1.50037412],
    [-2.05152671, 1.98626062, -0.59000837],
    [-0.78333082, -1.21731013, 0.69713417],
    [-1.95915437, -0.17725505, -0.97582275],
    [0.04164598, -0.47531991, -1.26098837],
    [-0.37343875, 0.4638171, -0.01383896],
    [-0.04278462, -0.59883687, -0.44309735]])


def test_epa_known_answer():
    # distance3d/test/test_epa.py:7-34
    cs = pack_colliders([C.ConvexHullVertices(EPA_VERTICES1), C.ConvexHullVertices(EPA_VERTICES2)])
    g = O.gjk_distance(cs, [[0, 1]])
    np.testing.assert_allclose(g["a"][0], g["b"][0], atol=1e-6)
    e = O.epa(cs, [[0, 1]], g["Y"])
    assert e["success"][0] == 1
    np.testing.assert_allclose(e["mtv"][0], [-0.387287, 0.179576, -0.176204], atol=5e-7)


def test_mpr_points_segments_spheres_known_answers():
    # distance3d/test/test_mpr.py:23-82
    p1 = C.ConvexHullVertices(np.array([[0.0, 0.0, 0.0]]))
    s1 = C.ConvexHullVertices(np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]))
    p2 = C.ConvexHullVertices(np.array([[1.0, 0.0, 0.0]]))
    cs = pack_colliders([
        p1, s1, p2, C.Sphere(np.zeros(3), 1.0), C.Sphere(np.zeros(3), 1.0),
        C.Sphere(np.array([0.0, 0.0, 1.0]), 0.5), C.Sphere(np.array([1.0, 0.0, 0.0]), 0.5)])
    r = O.mpr(cs, [[0, 0], [1, 1], [1, 0], [1, 2], [3, 4], [3, 5], [3, 6]])
    assert np.all(r["hit"] == 1)
    np.testing.assert_allclose(r["depth"], [0.0, 1.0, 0.0, 0.0, 2.0, 0.5, 0.5], atol=1e-12)
    np.testing.assert_allclose(r["dir"][0], 0, atol=1e-12)
    np.testing.assert_allclose(r["dir"][1], [-1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(r["pos"][1], [0.5, 0, 0], atol=1e-6)
    np.testing.assert_allclose(r["pos"][3], [1, 0, 0], atol=1e-6)
    assert abs(np.linalg.norm(r["dir"][4]) - 1.0) < 1e-12
    np.testing.assert_allclose(r["pos"][4], 0, atol=1e-6)
    np.testing.assert_allclose(r["dir"][5], [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(r["pos"][5], [0, 0, 0.75], atol=1e-6)
    np.testing.assert_allclose(r["dir"][6], [1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(r["pos"][6], [0.75, 0, 0], atol=1e-6)


def test_gjk_points_and_segments():
    # distance3d/test/test_gjk.py:7-53
    p = C.ConvexHullVertices(np.array([[0.0, 0.0, 0.0]]))
    r = dist(p, p)
    assert r["dist"][0] == 0.0
    q = C.ConvexHullVertices(np.array([[1.0, 0.0, 0.0]]))
    r = dist(p, q)
    assert r["dist"][0] == 1.0
    np.testing.assert_array_equal(r["a"][0], [0, 0, 0])
    np.testing.assert_array_equal(r["b"][0], [1, 0, 0])
    seg = C.ConvexHullVertices(np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]))
    seg2 = C.ConvexHullVertices(np.array([[0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]))
    assert dist(seg, seg2)["dist"][0] == 1.0


def test_gjk_spheres_analytic():
    # distance3d/test/test_gjk.py:132-160 style
    s1 = C.Sphere(np.zeros(3), 1.0)
    s2 = C.Sphere(np.array([0.0, 0.0, 3.0]), 1.0)
    r = dist(s1, s2)
    assert r["dist"][0] == pytest.approx(1.0, abs=1e-9)
    np.testing.assert_allclose(r["a"][0], [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(r["b"][0], [0, 0, 2], atol=1e-6)
    s3 = C.Sphere(np.array([0.0, 0.5, 0.0]), 1.0)
    assert dist(s1, s3)["dist"][0] == 0.0


def test_gjk_clipping():
    # distance3d/test/test_gjk_jolt.py:88-95
    s1 = C.Sphere(np.zeros(3), 1.0)
    s2 = C.Sphere(np.array([0.0, 0.0, 1000.0]), 1.0)
    r = dist(s1, s2, max_distance_squared=100000.0)
    assert r["status"][0] == 3 and r["dist"][0] == MAX_FLOAT


def test_margin_shifts_distance():
    # distance3d/test/test_gjk.py:476-482
    s1 = C.Sphere(np.zeros(3), 1.0)
    s2 = C.Sphere(np.array([0.0, 0.0, 4.0]), 1.0)
    d0 = dist(s1, s2)["dist"][0]
    d1 = dist(C.Margin(s1, 0.5), s2)["dist"][0]
    assert d0 - d1 == pytest.approx(0.5, abs=1e-9)


def test_containment_golden_values():
    # distance3d/test/test_containment.py:13-185
    cyl2origin = transform_from(
        R=active_matrix_from_extrinsic_euler_xyz([0.5, 0.3, 0.2]), p=np.array([0.1, 0.2, 0.3]))
    cs = pack_colliders([
        C.Sphere(np.array([0.1, 0.2, 0.3]), 0.5),
        C.Capsule(np.eye(4), 0.5, 1.0),
        C.Cylinder(cyl2origin, 0.5, 2.0),
    ])
    A = O.aabb(cs)
    np.testing.assert_allclose(A[0], [[-0.4, 0.6], [-0.3, 0.7], [-0.2, 0.8]], atol=1e-12)
    np.testing.assert_allclose(A[1], [[-0.5, 0.5], [-0.5, 0.5], [-1.0, 1.0]], atol=1e-12)
    # the cylinder box must contain the two cap centres +- 0 and be symmetric about the centre
    ctr = 0.5 * (A[2, :, 0] + A[2, :, 1])
    np.testing.assert_allclose(ctr, [0.1, 0.2, 0.3], atol=1e-12)
    axis = cyl2origin[:3, 2]
    assert np.all(A[2, :, 1] - ctr >= np.abs(axis) - 1e-12)
