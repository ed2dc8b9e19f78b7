"""Gilbert-Johnson-Keerthi queries, Jolt variant, batched on the GPU.

Drop-in names of the reference (distance3d/gjk/__init__.py:27-29):
``gjk`` = ``gjk_distance`` = ``gjk_distance_jolt`` (gjk/_gjk_jolt.py:138-221),
``gjk_intersection`` = ``gjk_intersection_jolt`` (gjk/_gjk_jolt.py:29-80), plus
``gjk_distance_jolt_iterations`` (gjk/_gjk_jolt.py:714-785).  The scalar calls
run a batch of one; ``gjk_distance_batch`` / ``gjk_intersection_batch`` take a
packed :class:`~distance3d_b200.pack.ColliderSet` and an int32 ``[P, 2]`` pair
index and return device tensors.
"""
import ctypes
from enum import Enum

import numpy as np

from . import _lib
from ._lib import MAX_FLOAT, c_dbl, c_i64, c_size, ptr
from .pack import pack_colliders

STATUS_NO_INTERSECTION = 0
STATUS_INTERSECTION = 1
STATUS_UNKNOWN = 2
STATUS_CLIPPED = 3
STATUS_SANITY_FAILED = 4
STATUS_MONOTONICITY = 5
STATUS_ITER_CAP = 6
STATUS_EPA_MAX_FACES = 7
STATUS_EPA_BAD_SIMPLEX = 8


class GjkState(Enum):
    NoIntersection = 0
    Intersection = 1
    Unknown = 2
    Clipped = 3


_workspaces = {}


def workspace(n_bytes, device, tag="gjk"):
    """Grow-only scratch buffer per (device, CUDA stream, tag): calls issued on different
    streams may run concurrently and must not share binning scratch or work counters; a
    buffer that is replaced stays alive until the work queued on its stream is done
    (record_stream)."""
    torch = _lib.torch_cuda()
    stream = torch.cuda.current_stream(device)
    key = (device.index, stream.cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < n_bytes:
        if buf is not None:
            buf.record_stream(stream)
        buf = torch.empty(int(n_bytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


class GjkResult:
    """Device-resident result of :func:`gjk_distance_batch`."""

    def __init__(self, dist, closest_a, closest_b, simplex, n_points, iters, status):
        self.dist = dist
        self.closest_a = closest_a
        self.closest_b = closest_b
        self.simplex = simplex
        self.n_points = n_points
        self.iters = iters
        self.status = status

    def cpu(self):
        return {k: (v.cpu().numpy() if v is not None else None) for k, v in self.__dict__.items()}


def gjk_distance_batch(colliders, pairs, tolerance=1e-10, max_distance_squared=100000.0,
                       sanity_check=1e-8, want_points=True, want_simplex=True,
                       want_iters=True, out=None, dtype="f64"):
    """Distance, closest points and simplex for every pair (device tensors).

    dtype="f64" (default) is the parity mode, bit-compatible with the reference.
    dtype="f32" opts into single-precision arithmetic: 99.9 % of the distances within
    1e-4 of the fp64 result for unit-scale shapes, rare early terminations up to 0.2 off
    (always an upper bound); tolerance >= 1e-6 and sanity_check >= 1e-3 are enforced.
    """
    if dtype not in ("f64", "f32"):
        raise ValueError("dtype must be 'f64' or 'f32'")
    if dtype == "f32":
        tolerance = max(tolerance, 1e-6)
        sanity_check = max(sanity_check, 1e-3)
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    pairs = _lib.as_device_pairs(pairs, dc.device)
    n = pairs.shape[0]
    f64 = dict(dtype=torch.float64, device=dc.device)
    i32 = dict(dtype=torch.int32, device=dc.device)
    if out is None:
        out = GjkResult(
            torch.empty(n, **f64),
            torch.empty((n, 3), **f64) if want_points else None,
            torch.empty((n, 3), **f64) if want_points else None,
            torch.empty((n, 4, 3), **f64) if want_simplex else None,
            torch.empty(n, **i32) if want_simplex else None,
            torch.empty(n, **i32) if want_iters else None,
            torch.empty(n, **i32))
    L = _lib.lib()
    ws_bytes = L.d3d_gjk_workspace_bytes(c_i64(n))
    ws = workspace(ws_bytes, dc.device)
    fn = L.d3d_gjk_distance if dtype == "f64" else L.d3d_gjk_distance_f32
    _lib._check(fn(
        ctypes.byref(dc.struct), ptr(pairs), c_i64(n), c_dbl(tolerance),
        c_dbl(max_distance_squared), c_dbl(sanity_check), ptr(out.dist), ptr(out.closest_a),
        ptr(out.closest_b), ptr(out.simplex), ptr(out.n_points), ptr(out.iters),
        ptr(out.status), ptr(ws), c_size(ws.numel()), _lib.stream_ptr()))
    return out


def gjk_intersection_batch(colliders, pairs, tolerance=1e-10, want_iters=False, dtype="f64"):
    """Boolean intersection for every pair: (hit uint8[P], iters|None, status int32[P]).
    dtype="f32": opt-in single-precision arithmetic (see gjk_distance_batch)."""
    if dtype not in ("f64", "f32"):
        raise ValueError("dtype must be 'f64' or 'f32'")
    if dtype == "f32":
        tolerance = max(tolerance, 1e-6)
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    pairs = _lib.as_device_pairs(pairs, dc.device)
    n = pairs.shape[0]
    hit = torch.empty(n, dtype=torch.uint8, device=dc.device)
    iters = torch.empty(n, dtype=torch.int32, device=dc.device) if want_iters else None
    status = torch.empty(n, dtype=torch.int32, device=dc.device)
    L = _lib.lib()
    ws_bytes = L.d3d_gjk_intersection_workspace_bytes(c_i64(n))
    ws = workspace(ws_bytes, dc.device)
    fn = L.d3d_gjk_intersection if dtype == "f64" else L.d3d_gjk_intersection_f32
    _lib._check(fn(
        ctypes.byref(dc.struct), ptr(pairs), c_i64(n), c_dbl(tolerance), ptr(hit), ptr(iters),
        ptr(status), ptr(ws), c_size(ws.numel()), _lib.stream_ptr()))
    return hit, iters, status


_PAIR01 = np.array([[0, 1]], dtype=np.int32)


def _raise_for_status(status):
    if status == STATUS_SANITY_FAILED:
        raise AssertionError("Sanity check failed")  # _gjk_jolt.py:216
    if status == STATUS_MONOTONICITY:
        raise AssertionError("prev_v_len_sq >= v_len_sq")  # _gjk_jolt.py:126,282
    if status == STATUS_ITER_CAP:
        raise RuntimeError("GJK did not terminate within the iteration cap")


def gjk_distance_jolt(collider1, collider2, tolerance=1e-10, max_distance_squared=100000.0,
                      sanity_check=1e-8):
    """Distance between two convex colliders (reference: _gjk_jolt.py:138-221).

    Returns ``(distance, closest_point1, closest_point2, simplex)`` or
    ``(MAX_FLOAT, None, None, None)`` when the pair was clipped.
    """
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    res = gjk_distance_batch(cs, _PAIR01, tolerance, max_distance_squared, sanity_check).cpu()
    cs.commit_mesh_state()
    status = int(res["status"][0])
    if status == STATUS_CLIPPED:
        return MAX_FLOAT, None, None, None
    _raise_for_status(status)
    return float(res["dist"][0]), res["closest_a"][0], res["closest_b"][0], res["simplex"][0]


def gjk_distance_jolt_iterations(collider1, collider2, tolerance=1e-10,
                                 max_distance_squared=100000.0):
    """Number of GJK iterations (reference: _gjk_jolt.py:714-785)."""
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    res = gjk_distance_batch(cs, _PAIR01, tolerance, max_distance_squared, float("inf")).cpu()
    cs.commit_mesh_state()
    return int(res["iters"][0])


def gjk_intersection_jolt(collider1, collider2, tolerance=1e-10):
    """Do two convex colliders intersect? (reference: _gjk_jolt.py:29-80)."""
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    hit, _, status = gjk_intersection_batch(cs, _PAIR01, tolerance)
    cs.commit_mesh_state()
    _raise_for_status(int(status[0]))
    return bool(hit[0])


def gjk_intersection_libccd_batch(colliders, pairs, max_iterations=100, want_iters=False, sort_by_type=True):
    """libccd-style boolean GJK for every pair (gjk/_gjk_libccd.py:14-266), one thread per pair:
    ``(hit uint8[P], iters int32[P] | None)``.  An independent algorithm for the same question as
    `gjk_intersection_batch`; the two agree except on touching configurations."""
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    pairs = _lib.as_device_pairs(pairs, dc.device)
    n = pairs.shape[0]
    hit = torch.empty(n, dtype=torch.uint8, device=dc.device)
    iters = torch.empty(n, dtype=torch.int32, device=dc.device) if want_iters else None
    perm = None
    if sort_by_type and n > 64:
        key = dc.type[pairs[:, 0].long()] * 16 + dc.type[pairs[:, 1].long()]
        perm = torch.argsort(key).to(torch.int32)
    _lib._check(_lib.lib().d3d_gjk_intersection_libccd(
        ctypes.byref(dc.struct), ptr(pairs), ptr(perm), c_i64(n), ctypes.c_int(max_iterations),
        ptr(hit), ptr(iters), _lib.stream_ptr()))
    return hit, iters


def gjk_intersection_libccd(collider1, collider2, max_iterations=100):
    """Do two convex colliders intersect? libccd variant (reference: _gjk_libccd.py:14-53)."""
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    hit, _ = gjk_intersection_libccd_batch(cs, _PAIR01, max_iterations)
    cs.commit_mesh_state()
    return bool(hit[0])


gjk = gjk_distance_jolt
gjk_distance = gjk_distance_jolt
gjk_intersection = gjk_intersection_jolt
