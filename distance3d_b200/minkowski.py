"""Minkowski-difference helpers (reference: distance3d/minkowski.py:5-74)."""
import numpy as np


class Simplex:
    """Simplex of Minkowski differences and the support points on both colliders."""

    def __init__(self):
        self.v = np.empty((4, 3))
        self.v1 = np.empty((4, 3))
        self.v2 = np.empty((4, 3))
        self.n_points = 0

    def __len__(self):
        return self.n_points

    def add_point(self, v, v1, v2):
        self.v[self.n_points] = v
        self.v1[self.n_points] = v1
        self.v2[self.n_points] = v2
        self.n_points += 1


def make_support_point(v1, v2):
    return v1 - v2, v1, v2


def support_function(collider1, collider2, search_direction):
    """Support point of A - B and the points on A and B (minkowski.py:23-50); both support
    maps are evaluated in one `d3d_support` launch."""
    from . import _lib
    from .pack import pack_colliders
    d = np.asarray(search_direction, dtype=float)
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    out = _lib.support(cs, np.array([0, 1], dtype=np.int32), np.stack((d, -d)))
    cs.commit_mesh_state()
    return make_support_point(out[0], out[1])


def minkowski_sum(vertices1, vertices2):
    """All pairwise sums of two vertex sets (minkowski.py:58-74)."""
    v1 = np.asarray(vertices1)
    v2 = np.asarray(vertices2)
    return (v1[:, np.newaxis, :] + v2[np.newaxis, :, :]).reshape(-1, v1.shape[1])
