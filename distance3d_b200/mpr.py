"""Minkowski portal refinement on the GPU (reference: distance3d/mpr.py:21-109)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_dbl, c_i64, c_int, ptr
from .pack import pack_colliders


def mpr_batch(colliders, pairs, mpr_tolerance=0.0001, max_iterations=100, penetration=True,
              sort_by_type=True):
    """MPR for many pairs; returns dict of device tensors hit, depth, dir, pos, status."""
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    pairs = _lib.as_device_pairs(pairs, dc.device)
    n = pairs.shape[0]
    f64 = dict(dtype=torch.float64, device=dc.device)
    out = dict(hit=torch.empty(n, dtype=torch.uint8, device=dc.device),
               status=torch.empty(n, dtype=torch.int32, device=dc.device),
               depth=torch.zeros(n, **f64) if penetration else None,
               dir=torch.zeros((n, 3), **f64) if penetration else None,
               pos=torch.zeros((n, 3), **f64) if penetration else None)
    perm = None
    if sort_by_type and n > 64:
        # group pairs by (typeA, typeB) so that warps evaluate one kind of support map
        key = dc.type[pairs[:, 0].long()] * 16 + dc.type[pairs[:, 1].long()]
        perm = torch.argsort(key).to(torch.int32)
    _lib._check(_lib.lib().d3d_mpr(
        ctypes.byref(dc.struct), ptr(pairs), ptr(perm), c_i64(n), c_dbl(mpr_tolerance),
        c_int(max_iterations), c_int(1 if penetration else 0), ptr(out["hit"]), ptr(out["depth"]),
        ptr(out["dir"]), ptr(out["pos"]), ptr(out["status"]), _lib.stream_ptr()))
    return out


_PAIR01 = np.array([[0, 1]], dtype=np.int32)


def mpr_intersection(collider1, collider2, mpr_tolerance=0.0001, max_iterations=100):
    """Intersection test with MPR (reference: mpr.py:21-50)."""
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    out = mpr_batch(cs, _PAIR01, mpr_tolerance, max_iterations, penetration=False)
    cs.commit_mesh_state()
    return bool(out["hit"][0])


def mpr_penetration(collider1, collider2, mpr_tolerance=0.0001, max_iterations=100):
    """MPR with penetration info (reference: mpr.py:53-109).

    Returns ``(intersection, depth, penetration_direction, contact_position)``;
    the last three are None when the colliders do not intersect.
    """
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    out = mpr_batch(cs, _PAIR01, mpr_tolerance, max_iterations, penetration=True)
    cs.commit_mesh_state()
    if not bool(out["hit"][0]):
        return False, None, None, None
    return (True, float(out["depth"][0]), out["dir"][0].cpu().numpy(), out["pos"][0].cpu().numpy())
