"""Broad phase + tetrahedron-pair intersection of the hydroelastic contact model.

The reference's `hydroelastic_contact.find_contact_surface` (_interface.py:52-101) is the other
consumer of `AabbTree.overlaps_aabb_tree` / `all_aabbs_overlap`: the tetrahedra of two rigid
bodies, expressed in one frame, go through the AABB broad phase and every candidate pair through
`intersect_tetrahedron_pair` (_tetrahedron_intersection.py:87-140) in a Python loop.  Here both
stages are batched on the device: boxes of all tetrahedra (`d3d_tetra_aabb`), an LBVH over the
second mesh queried with the boxes of the first (`Lbvh.overlap`), and one thread per candidate
pair for the contact plane and the contact polygon (`d3d_tetra_intersect_pairs`).  Pressure
integration (forces, wrenches), mesh generation and the `RigidBody` class are out of scope.
"""
import ctypes

import numpy as np

from . import _lib, aabb_tree
from ._lib import c_dbl, c_i64, c_int, ptr


def _dev(a, shape, dtype=None):
    torch = _lib.torch_cuda()
    dtype = dtype or torch.float64
    if isinstance(a, torch.Tensor):
        return a.to(device=torch.device("cuda", torch.cuda.current_device()), dtype=dtype).reshape(shape).contiguous()
    np_dtype = np.float64 if dtype == torch.float64 else np.int32
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype)).to(
        torch.device("cuda", torch.cuda.current_device())).reshape(shape)


def tetrahedral_mesh_aabbs(tetrahedra_points, device=False):
    """Boxes of tetrahedra, shape (n, 3, 2) (_mesh_processing.py:4-20)."""
    torch = _lib.torch_cuda()
    p = _dev(tetrahedra_points, (-1, 4, 3))
    out = torch.empty((p.shape[0], 3, 2), dtype=torch.float64, device=p.device)
    _lib._check(_lib.lib().d3d_tetra_aabb(ptr(p), c_i64(p.shape[0]), ptr(out), _lib.stream_ptr()))
    return out if device else out.cpu().numpy()


def barycentric_transforms(tetrahedra_points, device=False):
    """X with X.dot((r, 1)) = barycentric coordinates of r (_barycentric_transform.py:4-9)."""
    torch = _lib.torch_cuda()
    p = _dev(tetrahedra_points, (-1, 4, 3))
    out = torch.empty((p.shape[0], 4, 4), dtype=torch.float64, device=p.device)
    _lib._check(_lib.lib().d3d_tetra_barycentric(ptr(p), c_i64(p.shape[0]), ptr(out), _lib.stream_ptr()))
    return out if device else out.cpu().numpy()


class TetraPairResult:
    """Device tensors of :func:`intersect_tetrahedron_pairs_batch`."""

    def __init__(self, hit, plane, n_vertices, polygon, status):
        self.hit, self.plane, self.n_vertices, self.polygon, self.status = hit, plane, n_vertices, polygon, status

    def cpu(self):
        return {k: v.cpu().numpy() for k, v in self.__dict__.items()}


def intersect_tetrahedron_pairs_batch(pairs, tetrahedra_points1, tetrahedra_points2, epsilon1, epsilon2,
                                      X1=None, X2=None, youngs_modulus1=1.0, youngs_modulus2=1.0,
                                      max_vertices=12):
    """Contact plane and polygon of every candidate pair (device tensors)."""
    torch = _lib.torch_cuda()
    tp1, tp2 = _dev(tetrahedra_points1, (-1, 4, 3)), _dev(tetrahedra_points2, (-1, 4, 3))
    e1, e2 = _dev(epsilon1, (-1, 4)), _dev(epsilon2, (-1, 4))
    X1 = None if X1 is None else _dev(X1, (-1, 4, 4))
    X2 = None if X2 is None else _dev(X2, (-1, 4, 4))
    pairs = _lib.as_device_pairs(pairs, tp1.device)
    n = pairs.shape[0]
    dev = tp1.device
    res = TetraPairResult(torch.empty(n, dtype=torch.uint8, device=dev),
                          torch.empty((n, 4), dtype=torch.float64, device=dev),
                          torch.empty(n, dtype=torch.int32, device=dev),
                          torch.zeros((n, max_vertices, 3), dtype=torch.float64, device=dev),
                          torch.empty(n, dtype=torch.int32, device=dev))
    _lib._check(_lib.lib().d3d_tetra_intersect_pairs(
        ptr(pairs), c_i64(n), ptr(tp1), ptr(e1), ptr(X1), ptr(tp2), ptr(e2), ptr(X2),
        c_dbl(youngs_modulus1), c_dbl(youngs_modulus2), c_int(max_vertices), ptr(res.hit),
        ptr(res.plane), ptr(res.n_vertices), ptr(res.polygon), ptr(res.status), _lib.stream_ptr()))
    return res


def intersect_tetrahedron_pairs(pairs, tetrahedra_points1, tetrahedra_points2, epsilon1, epsilon2,
                                X1=None, X2=None, youngs_modulus1=1.0, youngs_modulus2=1.0):
    """Drop-in for _tetrahedron_intersection.py:7-84: ``(intersection, contact_planes,
    contact_polygons, intersecting_tetrahedra1, intersecting_tetrahedra2)``.  X1 / X2 may be the
    reference's dicts (tetrahedron index -> 4x4), arrays [n, 4, 4], or None (computed)."""
    def as_array(X, n):
        if X is None or not isinstance(X, dict):
            return X
        out = np.zeros((n, 4, 4))
        for k, v in X.items():
            out[int(k)] = v
        return out
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    if len(pairs) == 0:
        return False, [], [], [], []
    n1, n2 = len(tetrahedra_points1), len(tetrahedra_points2)
    res = intersect_tetrahedron_pairs_batch(pairs, tetrahedra_points1, tetrahedra_points2, epsilon1, epsilon2,
                                            as_array(X1, n1), as_array(X2, n2), youngs_modulus1,
                                            youngs_modulus2, max_vertices=24).cpu()
    sel = np.nonzero(res["hit"])[0]
    planes = res["plane"][sel] if len(sel) else []
    polygons = [res["polygon"][k, :res["n_vertices"][k]] for k in sel]
    return len(sel) > 0, planes, polygons, [int(i) for i in pairs[sel, 0]], [int(j) for j in pairs[sel, 1]]


def find_contact_pairs(tetrahedra_points1, epsilon1, tetrahedra_points2, epsilon2, youngs_modulus1=1.0,
                       youngs_modulus2=1.0, use_aabb_trees=True, max_vertices=12):
    """Broad + narrow phase of `find_contact_surface` (_interface.py:74-92) for two tetrahedral
    meshes given in ONE frame: returns ``(candidate pairs int32[C, 2] (device), TetraPairResult)``.
    use_aabb_trees=True: LBVH over mesh 2, queried with the boxes of mesh 1 (replaces
    `aabbtree_.overlaps_aabb_tree`); False: brute force (replaces `all_aabbs_overlap`)."""
    tp1, tp2 = _dev(tetrahedra_points1, (-1, 4, 3)), _dev(tetrahedra_points2, (-1, 4, 3))
    a1, a2 = tetrahedral_mesh_aabbs(tp1, device=True), tetrahedral_mesh_aabbs(tp2, device=True)
    if use_aabb_trees:
        pairs21, _ = aabb_tree.Lbvh(a2).overlap(a1, ordered=False, packet=False)   # (mesh 2, mesh 1)
        pairs = pairs21.flip(1).contiguous()
    else:
        pairs, _ = aabb_tree.brute_force_pairs(a1, a2)
    res = intersect_tetrahedron_pairs_batch(pairs, tp1, tp2, epsilon1, epsilon2, None, None,
                                            youngs_modulus1, youngs_modulus2, max_vertices)
    return pairs, res
