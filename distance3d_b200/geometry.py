"""Support functions and shape conversions with the reference's signatures
(distance3d/geometry.py:138-454), evaluated by the `d3d_support` / `d3d_prepare`
kernels on a batch of one."""
import numpy as np

from . import colliders as _c


def _f(a):
    return np.asarray(a, dtype=float)


def convert_box_to_vertices(box2origin, size):
    """8 world-frame vertices of a box (geometry.py:138-157)."""
    return _c.Box(_f(box2origin), _f(size)).vertices


def support_function_cylinder(search_direction, cylinder2origin, radius, length):
    return _c.Cylinder(_f(cylinder2origin), radius, length).support_function(search_direction)


def support_function_capsule(search_direction, capsule2origin, radius, height):
    return _c.Capsule(_f(capsule2origin), radius, height).support_function(search_direction)


def support_function_ellipsoid(search_direction, ellipsoid2origin, radii):
    return _c.Ellipsoid(_f(ellipsoid2origin), _f(radii)).support_function(search_direction)


def support_function_sphere(search_direction, center, radius):
    return _c.Sphere(_f(center), radius).support_function(search_direction)


def support_function_disk(search_direction, center, radius, normal):
    return _c.Disk(_f(center), radius, _f(normal)).support_function(search_direction)


def support_function_ellipse(search_direction, center, axes, radii):
    return _c.Ellipse(_f(center), _f(axes), _f(radii)).support_function(search_direction)


def support_function_cone(search_direction, cone2origin, radius, height):
    return _c.Cone(_f(cone2origin), radius, height).support_function(search_direction)
