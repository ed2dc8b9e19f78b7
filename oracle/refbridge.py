"""TEST INFRASTRUCTURE: import the real reference (build container only).

`/root/reference` is read-only and absent on the GPU box; this module is used
by oracle/validate_oracle.py and oracle/gen_golden.py, never by tests that run
on the GPU box, by smoke() or by bench.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get("D3D_REFERENCE", "/root/reference")


def setup():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    for p in (os.path.join(HERE, "refshim"), REPO, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)


def to_set(ref_colliders):
    """Reference collider objects -> distance3d_b200 ColliderSet (same numbers)."""
    setup()
    from distance3d_b200 import colliders as C
    from distance3d_b200.pack import pack_colliders
    out = []
    for c in ref_colliders:
        out.append(_convert(c, C))
    return pack_colliders(out)


def _convert(c, C):
    name = type(c).__name__
    if name == "Margin":
        return C.Margin(_convert(c.collider, C), c.margin)
    if name == "Box":
        return C.Box(np.array(c.box2origin), np.array(c.size))
    if name == "ConvexHullVertices":
        return C.ConvexHullVertices(np.array(c.vertices))
    if name == "MeshGraph":
        return C.MeshGraph(np.array(c.mesh2origin), np.array(c.vertices), np.array(c.triangles))
    if name == "Sphere":
        return C.Sphere(np.array(c.c), c.radius)
    if name == "Capsule":
        return C.Capsule(np.array(c.capsule2origin), c.radius, c.height)
    if name == "Ellipsoid":
        return C.Ellipsoid(np.array(c.ellipsoid2origin), np.array(c.radii))
    if name == "Cylinder":
        return C.Cylinder(np.array(c.cylinder2origin), c.radius, c.length)
    if name == "Disk":
        return C.Disk(np.array(c.c), c.radius, np.array(c.normal))
    if name == "Ellipse":
        return C.Ellipse(np.array(c.c), np.array(c.axes), np.array(c.radii))
    if name == "Cone":
        return C.Cone(np.array(c.cone2origin), c.radius, c.height)
    raise TypeError(name)


def random_reference_colliders(rs, n, names, hull_as_vertices=True, **kwargs):
    """n random reference colliders with the reference's own generators."""
    setup()
    from distance3d import colliders, random
    from distance3d.utils import transform_points
    out = []
    for _ in range(n):
        name = names[rs.randint(len(names))]
        args = random.RANDOM_GENERATORS[name](rs, **kwargs.get(name, {}))
        if name == "mesh" and hull_as_vertices:
            mesh2origin, vertices, triangles = args
            out.append(colliders.ConvexHullVertices(
                np.ascontiguousarray(transform_points(mesh2origin, vertices))))
        else:
            out.append(colliders.COLLIDERS[name](*args))
    return out
