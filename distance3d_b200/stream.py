"""Multi-buffered host -> device -> host pipeline for batches of GJK queries.

The narrow phase needs 336 bytes of input per pair in the reference's own layout (a 4x4
pose per collider) when every pair brings its own colliders, so a caller that holds its data
in HOST memory is PCIe bound.  `pin_batch(..., wire=True)` (the default) therefore ships the
compact wire records of `ColliderSet.wire()` - a sphere is 4 doubles, a capsule 14, 99 bytes
per collider on the primitive mix instead of 164 - and `d3d_unpack_colliders` expands them
into the structure of arrays in HBM (the record offsets are recomputed on the device from the
type bytes, `d3d_wire_offsets`, instead of travelling).  (The vertex pool travels only when the batch has hull /
mesh vertices in it: the 8 vertices of a box are derived data that `d3d_prepare` writes on
the device.)  This class
keeps `slots` sets of device buffers and CUDA streams: while the kernels of batch k
run, the inputs of batch k+1 are already on their way over PCIe and the results of
batch k-1 are copied back (H2D and D2H use different copy engines).  All work of one
batch is enqueued on the batch's stream through the C ABI (`stream` argument), nothing
synchronises the host until a result is collected.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_dbl, c_i64, c_size, ptr
from .pack import ColliderSet, DeviceColliders

_ARRAYS = ("type", "pose", "param", "vert_off", "vert_len", "verts")
# Poses travel as whole 4x4 matrices: a strided upload of rows 0-2 only (cudaMemcpy2DAsync, 96 of
# every 128 bytes) moves 19 % fewer bytes but measured 6 % slower end to end on B200 / PCIe 5.


def pin_batch(cs, pairs, wire=True, out=None):
    """Pinned host tensors of a ColliderSet and its pair list (staging copy).

    wire=True: compact wire records (type-specific, see `ColliderSet.wire`); wire=False: the
    structure-of-arrays layout as it is (4x4 poses).  `out`: a dict returned by an earlier call
    with the same sizes, whose pinned buffers are overwritten instead of allocated."""
    torch = _lib.torch_cuda()
    from .pack import HULL, MESH
    if cs.margin is not None or cs.graph_off is not None:
        raise NotImplementedError("GjkDistanceStream batches carry no Margin / MeshGraph fields; "
                                  "use gjk_distance_batch")
    if wire and out is not None:
        # pack straight into the pinned buffers of an earlier batch of the same shape
        cs.wire(out=(out["wire_type"].numpy(), out["wire_off"].numpy(), out["wire"].numpy()))
        if out.get("has_vertex_data", True):
            out["verts"].numpy()[...] = cs.verts
        out["pairs"].numpy()[...] = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        return out
    if wire:
        wt, wo, w = cs.wire()
        arrays = {"wire_type": wt, "wire_off": wo, "wire": w, "verts": cs.verts}
    else:
        arrays = {k: getattr(cs, k) for k in _ARRAYS}
    arrays["pairs"] = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    if out is not None:
        for k, a in arrays.items():
            out[k].numpy()[...] = a.reshape(out[k].shape)
        return out
    host = {k: torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for k, a in arrays.items()}
    host["wire_format"] = bool(wire)
    host["n_colliders"] = len(cs)
    # box slots of the pool are filled on the device; without hulls / meshes nothing is uploaded
    host["has_vertex_data"] = bool(np.any((cs.type == HULL) | (cs.type == MESH)))
    return host


class _Slot:
    def __init__(self, torch, dev, max_colliders, max_pairs, max_vertices):
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        n, p, m = max_colliders, max_pairs, max(max_vertices, 1)
        self.stream = torch.cuda.Stream(device=dev)
        self.dev = {"type": torch.empty(n, **i32), "pose": torch.empty((n, 4, 4), **f64),
                    "param": torch.empty((n, 3), **f64), "vert_off": torch.empty(n, **i32),
                    "vert_len": torch.empty(n, **i32), "verts": torch.empty((m, 3), **f64),
                    "pairs": torch.empty((p, 2), **i32),
                    # wire records: at most 16 doubles per collider
                    "wire_type": torch.empty(n, dtype=torch.uint8, device=dev),
                    "wire_off": torch.empty(n, **i32), "wire": torch.empty(16 * n, **f64)}
        self.out = {"dist": torch.empty(p, **f64), "closest_a": torch.empty((p, 3), **f64),
                    "closest_b": torch.empty((p, 3), **f64), "status": torch.empty(p, **i32)}
        self.host_out = {k: torch.empty_like(v, device="cpu").pin_memory() for k, v in self.out.items()}
        L = _lib.lib()
        self.ws = torch.empty(L.d3d_gjk_workspace_bytes(c_i64(p)), dtype=torch.uint8, device=dev)
        self.done = torch.cuda.Event()
        self.busy = False
        self.n_pairs = 0


class GjkDistanceStream:
    """Pipelined `gjk_distance` over host-resident batches.

    Parameters: capacity of one batch (colliders, pairs, pool vertices) and the number
    of slots in flight (2 = double buffering).
    """

    def __init__(self, max_colliders, max_pairs, max_vertices=1, slots=2, device=None,
                 tolerance=1e-10, max_distance_squared=100000.0, sanity_check=1e-8):
        torch = _lib.torch_cuda()
        self.torch = torch
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self.slots = [_Slot(torch, self.device, max_colliders, max_pairs, max_vertices)
                      for _ in range(slots)]
        self.params = (tolerance, max_distance_squared, sanity_check)
        self._next = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host):
        """Enqueue one batch (dict from :func:`pin_batch`); returns the slot index (ticket)."""
        torch = self.torch
        idx = self._next
        self._next = (self._next + 1) % len(self.slots)
        slot = self.slots[idx]
        if slot.busy:
            raise RuntimeError("slot still holds an uncollected result; call result() first")
        wire = host.get("wire_format", False)
        n = host["n_colliders"] if wire else host["type"].shape[0]
        p = host["pairs"].shape[0]
        m = host["verts"].shape[0]
        if (n > slot.dev["type"].shape[0] or p > slot.dev["pairs"].shape[0]
                or m > slot.dev["verts"].shape[0]):
            raise ValueError("batch (%d colliders, %d pairs, %d vertices) exceeds the stream's capacity"
                             % (n, p, m))
        L = _lib.lib()
        h2d = 0
        with torch.cuda.stream(slot.stream):
            views = {}
            for k in _ARRAYS + ("pairs",):
                size = {"verts": m, "pairs": p}.get(k, n)
                views[k] = slot.dev[k][:size]
            # the record offsets are recomputed on the device from the types (d3d_wire_offsets)
            sent = ("wire_type", "wire", "verts", "pairs") if wire else _ARRAYS + ("pairs",)
            for k in sent:
                if k == "verts" and not host.get("has_vertex_data", True):
                    continue
                dst = views[k] if k in views else slot.dev[k][:host[k].shape[0]]
                dst.copy_(host[k].reshape(dst.shape), non_blocking=True)
                h2d += host[k].numel() * host[k].element_size()
            if wire:   # expand the records into the structure of arrays (HBM to HBM)
                _lib._check(L.d3d_wire_offsets(ptr(slot.dev["wire_type"]), c_i64(n), ptr(slot.dev["wire_off"]),
                                               ctypes.c_void_p(slot.stream.cuda_stream)))
                _lib._check(L.d3d_unpack_colliders(
                    ptr(slot.dev["wire_type"]), ptr(slot.dev["wire_off"]), ptr(slot.dev["wire"]), c_i64(n),
                    ptr(views["type"]), ptr(views["pose"]), ptr(views["param"]), ptr(views["vert_off"]),
                    ptr(views["vert_len"]), ctypes.c_void_p(slot.stream.cuda_stream)))
            dc = DeviceColliders.__new__(DeviceColliders)
            dc._init(views["type"], views["pose"], views["param"], views["vert_off"],
                     views["vert_len"], views["verts"], None)     # d3d_prepare on this stream
            s = ctypes.c_void_p(slot.stream.cuda_stream)
            tol, mds, sc = self.params
            _lib._check(L.d3d_gjk_distance(
                ctypes.byref(dc.struct), ptr(views["pairs"]), c_i64(p), c_dbl(tol), c_dbl(mds), c_dbl(sc),
                ptr(slot.out["dist"]), ptr(slot.out["closest_a"]), ptr(slot.out["closest_b"]), None,
                None, None, ptr(slot.out["status"]), ptr(slot.ws), c_size(slot.ws.numel()), s))
            d2h = 0
            for k, v in slot.out.items():
                slot.host_out[k][:p].copy_(v[:p], non_blocking=True)
                d2h += v[:p].numel() * v.element_size()
            slot.done.record(slot.stream)
        slot.busy = True
        slot.n_pairs = p
        slot.keep = (dc, views)   # keep the views alive until the result is collected
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        return idx

    def result(self, ticket):
        """Wait for a batch and return its pinned host results (valid until the slot is reused)."""
        slot = self.slots[ticket]
        slot.done.synchronize()
        slot.busy = False
        p = slot.n_pairs
        return {k: v[:p] for k, v in slot.host_out.items()}

    def drain_into(self, stream):
        """Make `stream` wait for everything enqueued so far (for event timing)."""
        for slot in self.slots:
            stream.wait_stream(slot.stream)
