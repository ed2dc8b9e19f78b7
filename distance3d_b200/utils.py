"""Math helpers with the reference's names (distance3d/utils.py:7-230, host side, numpy).

Only the helpers that user code of the hot path touches are provided; they are plain
data plumbing (pose composition, normalisation of inputs), not part of the accelerated path.
"""
import numpy as np

from ._transforms import plane_basis_from_normal, invert_transform  # noqa: F401

MAX_FLOAT = np.finfo(float).max
EPSILON = np.finfo(float).eps
HALF_PI = 0.5 * np.pi


def norm_vector(v):
    """Unit vector, or the unchanged input when its norm is zero (utils.py:12-30)."""
    norm = np.linalg.norm(v)
    if norm == 0.0:
        return v
    return v / norm


def scalar_triple_product(a, b, c):
    """a . (b x c) (utils.py:33-73)."""
    return np.dot(a, np.cross(b, c))


def transform_point(A2B, point_in_A):
    """Point from frame A to frame B (utils.py:125-143)."""
    return A2B[:3, 3] + np.dot(A2B[:3, :3], point_in_A)


def transform_points(A2B, points_in_A):
    """Points from frame A to frame B (utils.py:146-164)."""
    return np.dot(points_in_A, A2B[:3, :3].T) + A2B[:3, 3]


def transform_directions(A2B, directions_in_A):
    """Directions from frame A to frame B (utils.py:167-185)."""
    return np.dot(directions_in_A, A2B[:3, :3].T)


def inverse_transform_point(A2B, point_in_B):
    """Point from frame B to frame A (utils.py:188-207)."""
    RT = A2B[:3, :3].T
    return np.dot(RT, point_in_B) - np.dot(RT, A2B[:3, 3])
