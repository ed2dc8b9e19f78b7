"""Bounding volumes (reference: distance3d/containment.py:6-229), computed by the
`d3d_aabb` kernel on a batch of one.  Each function returns ``(mins, maxs)``."""
import numpy as np

from . import _lib, colliders as _c


def _mins_maxs(collider):
    box = _lib.aabb(collider._set())[0]
    return box[:, 0].copy(), box[:, 1].copy()


def axis_aligned_bounding_box(P):
    return _mins_maxs(_c.ConvexHullVertices(np.asarray(P, dtype=float)))


def sphere_aabb(center, radius):
    return _mins_maxs(_c.Sphere(np.asarray(center, dtype=float), radius))


def box_aabb(box2origin, size):
    return _mins_maxs(_c.Box(np.asarray(box2origin, dtype=float), np.asarray(size, dtype=float)))


def cylinder_aabb(cylinder2origin, radius, length):
    return _mins_maxs(_c.Cylinder(np.asarray(cylinder2origin, dtype=float), radius, length))


def capsule_aabb(capsule2origin, radius, height):
    return _mins_maxs(_c.Capsule(np.asarray(capsule2origin, dtype=float), radius, height))


def ellipsoid_aabb(ellipsoid2origin, radii):
    return _mins_maxs(_c.Ellipsoid(np.asarray(ellipsoid2origin, dtype=float),
                                   np.asarray(radii, dtype=float)))


def disk_aabb(center, radius, normal):
    return _mins_maxs(_c.Disk(np.asarray(center, dtype=float), radius,
                              np.asarray(normal, dtype=float)))


def cone_aabb(cone2origin, radius, height):
    return _mins_maxs(_c.Cone(np.asarray(cone2origin, dtype=float), radius, height))


def ellipse_aabb(center, axes, radii):
    return _mins_maxs(_c.Ellipse(np.asarray(center, dtype=float), np.asarray(axes, dtype=float),
                                 np.asarray(radii, dtype=float)))
