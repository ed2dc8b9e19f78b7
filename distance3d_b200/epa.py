"""Expanding polytope algorithm on the GPU (reference: distance3d/epa.py:9-78).

``epa(simplex, collider1, collider2, ...)`` keeps the reference's signature and
return value ``(mtv, faces, success)``; ``epa_batch`` takes a packed collider set,
a pair index and the GJK simplices of the pairs (`GjkResult.simplex`).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_dbl, c_i64, c_int, c_size, ptr
from .gjk import STATUS_EPA_MAX_FACES, workspace
from .pack import pack_colliders


class EpaResult:
    def __init__(self, mtv, success, n_faces, iters, status, faces=None, deferred=None):
        self.mtv = mtv
        self.success = success
        self.n_faces = n_faces
        self.iters = iters
        self.status = status
        self.faces = faces
        # int32[1]: pairs the thread-per-pair kernel handed to the warp kernel (diagnostic)
        self.deferred = deferred

    def cpu(self):
        return {k: (v.cpu().numpy() if v is not None else None) for k, v in self.__dict__.items()}


def epa_batch(colliders, pairs, simplices, max_iter=64, max_loose_edges=32, max_faces=64,
              epsilon=1e-8, want_faces=False, n_points=None):
    """Minimum translation vectors for many intersecting pairs (device tensors).

    `simplices` f64[P,4,3]: GJK simplex of each pair; EPA is defined only where GJK
    ended with a full 4-point simplex.  status = 7 where the
    reference would raise its `max_faces` AssertionError (epa.py:128).  `n_points` int32[P]
    (optional, GjkResult.n_points): pairs with fewer than 4 simplex points are skipped with
    status = 8 instead of running on rows GJK never wrote.
    """
    torch = _lib.torch_cuda()
    dc = _lib.as_device_colliders(colliders)
    pairs = _lib.as_device_pairs(pairs, dc.device)
    n = pairs.shape[0]
    if isinstance(simplices, torch.Tensor):
        Y = simplices.to(device=dc.device, dtype=torch.float64).reshape(n, 4, 3).contiguous()
    else:
        Y = torch.from_numpy(np.ascontiguousarray(simplices, dtype=np.float64)).to(dc.device).reshape(n, 4, 3)
    f64 = dict(dtype=torch.float64, device=dc.device)
    i32 = dict(dtype=torch.int32, device=dc.device)
    res = EpaResult(torch.empty((n, 3), **f64), torch.empty(n, dtype=torch.uint8, device=dc.device),
                    torch.empty(n, **i32), torch.empty(n, **i32), torch.empty(n, **i32),
                    torch.empty((n, max_faces, 4, 3), **f64) if want_faces else None)
    if n_points is not None:
        n_points = n_points.to(device=dc.device, dtype=torch.int32).contiguous()
    L = _lib.lib()
    ws = workspace(L.d3d_epa_workspace_bytes(c_i64(n)), dc.device, "epa")
    _lib._check(L.d3d_epa(
        ctypes.byref(dc.struct), ptr(pairs), c_i64(n), ptr(Y), ptr(n_points), c_int(max_iter),
        c_int(max_loose_edges), c_int(max_faces), c_dbl(epsilon), ptr(res.mtv), ptr(res.success),
        ptr(res.n_faces), ptr(res.iters), ptr(res.status), ptr(res.faces), ptr(ws),
        c_size(ws.numel()), _lib.stream_ptr()))
    res.deferred = ws[8:12].view(torch.int32).clone()
    return res


def epa(simplex, collider1, collider2, max_iter=64, max_loose_edges=32, max_faces=64,
        epsilon=1e-8):
    """Expanding Polytope Algorithm (reference: epa.py:9-78).

    Returns ``(mtv, faces, success)``: the minimum translation vector to be added
    to the second collider (or subtracted from the first), the polytope faces
    ``(n_faces, 4, 3)`` and whether EPA converged.
    """
    cs = pack_colliders([collider1, collider2], track_mesh_state=True)
    res = epa_batch(cs, np.array([[0, 1]], dtype=np.int32),
                    np.asarray(simplex, dtype=np.float64).reshape(1, 4, 3), max_iter,
                    max_loose_edges, max_faces, epsilon, want_faces=True).cpu()
    cs.commit_mesh_state()
    if int(res["status"][0]) == STATUS_EPA_MAX_FACES:
        raise AssertionError("self.n_faces < self.max_faces")  # epa.py:128
    n_faces = int(res["n_faces"][0])
    return res["mtv"][0], res["faces"][0, :n_faces], bool(res["success"][0])
