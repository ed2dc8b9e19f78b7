"""Host-side rigid-transform helpers (numpy).

These restate the handful of pytransform3d functions that the reference's
input generators and scene glue rely on (reference call sites:
distance3d/random.py:166,222-450, distance3d/broad_phase.py:44,111,148).
pytransform3d is an un-vendored dependency of the reference (setup.py:26, no
pinned version); only the *distribution* of the generated poses matters for
parity because both the oracle and the CUDA path consume the same arrays.
"""
import math

import numpy as np


def norm_vector(v):
    """Unit vector or the unchanged input when its norm is 0."""
    n = np.linalg.norm(v)
    if n == 0.0:
        return v
    return np.asarray(v) / n


def perpendicular_to_vector(a):
    """Some vector perpendicular to a (pytransform3d.rotations semantics)."""
    a = np.asarray(a, dtype=float)
    if abs(a[1]) < 1e-12 and abs(a[2]) < 1e-12:
        return np.array([0.0, 1.0, 0.0]) if abs(a[0]) > 1e-12 else np.array([1.0, 0.0, 0.0])
    # cross(a, unit_x)
    return np.array([0.0, a[2], -a[1]])


def matrix_from_axis_angle(axis, angle):
    """Active rotation matrix from unit axis and angle (Rodrigues)."""
    ux, uy, uz = axis
    c = math.cos(angle)
    s = math.sin(angle)
    ci = 1.0 - c
    return np.array([
        [ci * ux * ux + c, ci * ux * uy - uz * s, ci * ux * uz + uy * s],
        [ci * uy * ux + uz * s, ci * uy * uy + c, ci * uy * uz - ux * s],
        [ci * uz * ux - uy * s, ci * uz * uy + ux * s, ci * uz * uz + c]])


def active_matrix_from_angle(basis, angle):
    axis = np.zeros(3)
    axis[basis] = 1.0
    return matrix_from_axis_angle(axis, angle)


def active_matrix_from_extrinsic_euler_xyz(e):
    """Rz(e[2]) Ry(e[1]) Rx(e[0]) (extrinsic x-y-z = URDF rpy)."""
    return active_matrix_from_angle(2, e[2]).dot(
        active_matrix_from_angle(1, e[1])).dot(
        active_matrix_from_angle(0, e[0]))


def transform_from(R, p):
    A2B = np.eye(4)
    A2B[:3, :3] = R
    A2B[:3, 3] = p
    return A2B


def transform_from_exponential_coordinates(Stheta):
    """SE(3) exponential map of a 6-vector (omega*theta, v*theta)."""
    Stheta = np.asarray(Stheta, dtype=float)
    theta = np.linalg.norm(Stheta[:3])
    if theta == 0.0:
        return transform_from(np.eye(3), Stheta[3:])
    w = Stheta[:3] / theta
    v = Stheta[3:] / theta
    R = matrix_from_axis_angle(w, theta)
    W = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    V = (np.eye(3) * theta + (1.0 - math.cos(theta)) * W
         + (theta - math.sin(theta)) * W.dot(W))
    return transform_from(R, V.dot(v))


def random_transform(rng, mean=None, cov=None):
    """Random pose: exp of a 6-D standard normal sample, left-applied to mean."""
    if cov is None:
        sample = rng.randn(6) if hasattr(rng, "randn") else rng.standard_normal(6)
    else:
        sample = rng.multivariate_normal(mean=np.zeros(6), cov=cov)
    delta = transform_from_exponential_coordinates(sample)
    if mean is None:
        return delta
    return np.dot(delta, mean)


def concat(A2B, B2C):
    return np.dot(B2C, A2B)


def invert_transform(A2B):
    B2A = np.eye(4)
    RT = A2B[:3, :3].T
    B2A[:3, :3] = RT
    B2A[:3, 3] = -RT.dot(A2B[:3, 3])
    return B2A


def assert_transform(A2B, *args, **kwargs):
    A2B = np.asarray(A2B)
    assert A2B.shape == (4, 4)
    np.testing.assert_allclose(A2B[:3, :3].dot(A2B[:3, :3].T), np.eye(3), atol=1e-6)
    np.testing.assert_allclose(A2B[3], [0, 0, 0, 1], atol=1e-12)


def plane_basis_from_normal(plane_normal):
    """Two unit vectors spanning the plane with the given unit normal.

    Same deterministic choice as the reference (distance3d/utils.py:78-122):
    the larger of |nx|, |ny| decides which axis is eliminated.
    """
    n = plane_normal
    if abs(n[0]) >= abs(n[1]):
        length = math.sqrt(n[0] * n[0] + n[2] * n[2])
        x = np.array([-n[2] / length, 0.0, n[0] / length])
        y = np.array([n[1] * x[2], n[2] * x[0] - n[0] * x[2], -n[1] * x[0]])
    else:
        length = math.sqrt(n[1] * n[1] + n[2] * n[2])
        x = np.array([0.0, n[2] / length, -n[1] / length])
        y = np.array([n[1] * x[2] - n[2] * x[1], -n[0] * x[2], n[0] * x[1]])
    return x, y
