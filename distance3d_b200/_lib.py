"""ctypes binding of libd3d_b200.so (the C ABI in include/d3d_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is available, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# D3D_B200_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("D3D_B200_LIB") or os.path.join(_HERE, "libd3d_b200.so")

MAX_FLOAT = float(np.finfo(float).max)
EPSILON = float(np.finfo(float).eps)

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p
c_size = ctypes.c_size_t

_lib = None


class D3DError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise D3DError(
                "distance3d_b200: %s is missing -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)"
                % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.d3d_last_error_string.restype = ctypes.c_char_p
        for name in ("d3d_gjk_workspace_bytes", "d3d_gjk_intersection_workspace_bytes"):
            getattr(L, name).restype = c_size
            getattr(L, name).argtypes = [c_i64]
        for name in ("d3d_epa_workspace_bytes", "d3d_bvh_workspace_bytes",
                     "d3d_bvh_query_workspace_bytes"):
            if hasattr(L, name):
                getattr(L, name).restype = c_size
                getattr(L, name).argtypes = [c_i64]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise D3DError(lib().d3d_last_error_string().decode())


def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        raise D3DError("distance3d_b200 needs a CUDA device (no CPU fallback exists)")
    return torch


def stream_ptr():
    torch = torch_cuda()
    return c_ptr(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_ptr(t.data_ptr()) if t is not None else None


def as_device_colliders(colliders):
    """ColliderSet | DeviceColliders -> DeviceColliders on the current device."""
    from .pack import ColliderSet, DeviceColliders, pack_colliders
    torch_cuda()
    if isinstance(colliders, DeviceColliders):
        return colliders
    if isinstance(colliders, ColliderSet):
        return colliders.device()
    return pack_colliders(list(colliders)).device()


def as_device_pairs(pairs, device):
    torch = torch_cuda()
    if isinstance(pairs, torch.Tensor):
        t = pairs.to(device=device, dtype=torch.int32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(pairs, dtype=np.int32)).to(device)
    return t.reshape(-1, 2).contiguous()


# ---------------------------------------------------------------------------
def prepare(dc):
    """Generate box vertices in the device pool (d3d_prepare)."""
    _check(lib().d3d_prepare(ctypes.byref(dc.struct), ptr(dc.verts), stream_ptr()))


def support(colliders, idx, dirs):
    torch = torch_cuda()
    dc = as_device_colliders(colliders)
    idx_t = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(dc.device)
    dirs_t = torch.from_numpy(np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)).to(dc.device)
    out = torch.empty((len(idx_t), 3), dtype=torch.float64, device=dc.device)
    _check(lib().d3d_support(ctypes.byref(dc.struct), ptr(idx_t), ptr(dirs_t), c_i64(len(idx_t)),
                             ptr(out), stream_ptr()))
    return out.cpu().numpy()


def center(colliders):
    torch = torch_cuda()
    dc = as_device_colliders(colliders)
    out = torch.empty((dc.n, 3), dtype=torch.float64, device=dc.device)
    _check(lib().d3d_center(ctypes.byref(dc.struct), ptr(out), stream_ptr()))
    return out.cpu().numpy()


def aabb_device(colliders):
    torch = torch_cuda()
    dc = as_device_colliders(colliders)
    out = torch.empty((dc.n, 3, 2), dtype=torch.float64, device=dc.device)
    _check(lib().d3d_aabb(ctypes.byref(dc.struct), ptr(out), stream_ptr()))
    return out


def aabb(colliders):
    return aabb_device(colliders).cpu().numpy()


def box_vertices(colliders):
    """World-frame vertices of the boxes of a set: array [n_boxes, 8, 3]."""
    from .pack import BOX
    dc = as_device_colliders(colliders)
    verts = dc.verts.cpu().numpy()
    types = dc.type.cpu().numpy()
    offs = dc.vert_off.cpu().numpy()
    return np.array([verts[o:o + 8] for t, o in zip(types, offs) if t == BOX])


def debug_norm(v, mode=0):
    """Test hook: device emulation of the reference's BLAS dnrm2 on rows of v[n,3]."""
    torch = torch_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    v_t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)).to(dev)
    out = torch.empty(len(v_t), dtype=torch.float64, device=dev)
    _check(lib().d3d_debug_norm(ptr(v_t), c_i64(len(v_t)), ptr(out), c_int(mode), stream_ptr()))
    return out.cpu().numpy()


def debug_vdiv(v, s):
    """Test hook: rows of v[n,3] divided by s[n] with the device's vector division."""
    torch = torch_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    v_t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)).to(dev)
    s_t = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64).reshape(-1)).to(dev)
    out = torch.empty_like(v_t)
    _check(lib().d3d_debug_vdiv(ptr(v_t), ptr(s_t), c_i64(len(v_t)), ptr(out), stream_ptr()))
    return out.cpu().numpy()
