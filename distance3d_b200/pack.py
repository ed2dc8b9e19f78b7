"""Structure-of-arrays packing of colliders (host side).

The reference keeps one Python object per collider (distance3d/colliders.py:17-646)
and calls its `support_function` thousands of times per pair.  Here a set of N
colliders is ONE record of flat arrays (see include/d3d_types.h) that lives in
HBM and is indexed by the kernels:

    type      int32[N]     type tag
    pose      f64[N,4,4]   row-major collider2origin
    param     f64[N,3]     radius / height / size / radii ...
    vert_off  int32[N]     range into the shared vertex pool
    vert_len  int32[N]
    verts     f64[M,3]     box: 8 world vertices (written on the device by
                           d3d_prepare); hull: world frame; mesh: local frame
    margin    f64[N]       optional Margin wrapper
    graph_off int32[N]     MeshGraph: offset of the adjacency record in `graph` (-1: none)
    graph     int32[G]     adjacency records (distance3d_b200.mesh.build_mesh_graph)
    mesh_start int32[N]    MeshGraph: cached start vertex of the object (-1: record default)

`ColliderSet` holds the host copy (numpy) and, lazily, the device copy (torch
tensors on the current CUDA device).
"""
import ctypes

import numpy as np

SPHERE, CAPSULE, BOX, ELLIPSOID, CYLINDER, HULL, MESH, DISK, ELLIPSE, CONE = range(10)
D3D_NUM_TYPES = 10
TYPE_NAMES = ["sphere", "capsule", "box", "ellipsoid", "cylinder", "hull", "mesh",
              "disk", "ellipse", "cone"]


class CColliders(ctypes.Structure):
    """ctypes mirror of `struct d3d_colliders` (include/d3d_types.h)."""
    _fields_ = [
        ("n", ctypes.c_int64),
        ("type", ctypes.c_void_p),
        ("pose", ctypes.c_void_p),
        ("param", ctypes.c_void_p),
        ("vert_off", ctypes.c_void_p),
        ("vert_len", ctypes.c_void_p),
        ("verts", ctypes.c_void_p),
        ("margin", ctypes.c_void_p),
        ("graph_off", ctypes.c_void_p),
        ("graph", ctypes.c_void_p),
        ("mesh_start", ctypes.c_void_p),
        ("mesh_last", ctypes.c_void_p),
    ]


class ColliderSet:
    """Packed set of colliders.

    Build it with :func:`pack_colliders` (from collider objects) or
    :meth:`from_arrays`.  Host arrays are the source of truth until
    :meth:`device` is called; device tensors are cached per CUDA device.
    """

    def __init__(self, type_, pose, param, vert_off, vert_len, verts, margin=None,
                 boxes_prepared=False, graph_off=None, graph=None, mesh_start=None):
        n = len(type_)
        self.type = np.ascontiguousarray(type_, dtype=np.int32)
        self.pose = np.ascontiguousarray(pose, dtype=np.float64).reshape(n, 4, 4)
        self.param = np.ascontiguousarray(param, dtype=np.float64).reshape(n, 3)
        self.vert_off = np.ascontiguousarray(vert_off, dtype=np.int32)
        self.vert_len = np.ascontiguousarray(vert_len, dtype=np.int32)
        self.verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 3)
        if len(self.verts) == 0:
            self.verts = np.zeros((1, 3))  # never hand out a null pool pointer
        self.margin = None if margin is None else np.ascontiguousarray(margin, dtype=np.float64)
        self.boxes_prepared = boxes_prepared
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        self.graph_off, self.graph, self.mesh_start = i32(graph_off), i32(graph), i32(mesh_start)
        if self.graph_off is None:
            self.graph = None
        hv = (self.type == HULL) | (self.type == MESH)
        if np.any(hv & (self.vert_len <= 0)):
            raise ValueError("ConvexHullVertices / MeshGraph colliders need at least one vertex")
        # scalar API: the MeshGraph objects whose cached start vertex follows the calls
        self.mesh_objects = None
        self._device = {}

    from_arrays = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def __len__(self):
        return len(self.type)

    @property
    def n_vertices(self):
        return len(self.verts)

    def host_struct(self):
        """`d3d_colliders` over the HOST arrays (for the CPU oracle in tests)."""
        s = CColliders()
        s.n = len(self)
        s.type = self.type.ctypes.data
        s.pose = self.pose.ctypes.data
        s.param = self.param.ctypes.data
        s.vert_off = self.vert_off.ctypes.data
        s.vert_len = self.vert_len.ctypes.data
        s.verts = self.verts.ctypes.data
        s.margin = None if self.margin is None else self.margin.ctypes.data
        s.graph_off = None if self.graph_off is None else self.graph_off.ctypes.data
        s.graph = None if self.graph is None else self.graph.ctypes.data
        s.mesh_start = None if self.mesh_start is None else self.mesh_start.ctypes.data
        s.mesh_last = None
        return s

    def device(self, device=None):
        """Upload (once) and return the device-resident buffers."""
        import torch
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        key = (device.type, device.index)
        if key not in self._device:
            self._device[key] = DeviceColliders(self, device)
        return self._device[key]

    def invalidate_device(self):
        """Call after editing the host arrays in place (poses, parameters)."""
        self._device = {}
        self.boxes_prepared = False

    def subset(self, idx):
        """New set with the colliders `idx` (vertex pool is re-packed)."""
        idx = np.asarray(idx)
        vert_len = self.vert_len[idx]
        vert_off = np.zeros(len(idx), dtype=np.int64)
        if len(idx):
            vert_off[1:] = np.cumsum(vert_len[:-1])
        total = int(vert_len.sum())
        verts = np.zeros((total, 3))
        for k, i in enumerate(idx):
            o, l = int(self.vert_off[i]), int(self.vert_len[i])
            verts[vert_off[k]:vert_off[k] + l] = self.verts[o:o + l]
        return ColliderSet(self.type[idx], self.pose[idx], self.param[idx], vert_off, vert_len,
                           verts, None if self.margin is None else self.margin[idx],
                           boxes_prepared=self.boxes_prepared,
                           graph_off=None if self.graph_off is None else self.graph_off[idx],
                           graph=self.graph,
                           mesh_start=None if self.mesh_start is None else self.mesh_start[idx])

    def wire(self, out=None, n_threads=None):
        """Compact wire records (include/d3d_b200.h d3d_unpack_colliders): ``(wire_type uint8[n],
        wire_off int32[n], wire f64[total])``, packed by the library's host-side packer
        (`d3d_pack_wire_host`, multi-threaded C++).  `out`: three arrays to fill in place (e.g.
        views of pinned buffers)."""
        import os
        from . import _lib
        L = _lib.lib()
        L.d3d_wire_size.restype = ctypes.c_int64
        n = len(self)
        total = int(L.d3d_wire_size(ctypes.c_void_p(self.type.ctypes.data), ctypes.c_int64(n)))
        if total < 0:
            raise _lib.D3DError(L.d3d_last_error_string().decode())
        if out is None:
            out = (np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.int32), np.empty(max(total, 1)))
        wt, wo, w = out
        if len(wt) < n or len(wo) < n or len(w) < total:
            raise ValueError("wire output buffers are too small")
        s = self.host_struct()
        if n_threads is None:
            try:
                n_threads = len(os.sched_getaffinity(0))
            except AttributeError:
                n_threads = os.cpu_count() or 1
        _lib._check(L.d3d_pack_wire_host(ctypes.byref(s), ctypes.c_void_p(wt.ctypes.data),
                                         ctypes.c_void_p(wo.ctypes.data), ctypes.c_void_p(w.ctypes.data),
                                         ctypes.c_int(int(n_threads))))
        return wt[:n], wo[:n], w[:max(total, 1)]

    def commit_mesh_state(self, device=None):
        """Scalar API: copy the vertex every MeshGraph ended on back into the objects
        (`first_idx` caching of the reference, mesh.py:85)."""
        if not self.mesh_objects:
            return
        last = self.device(device).mesh_last.cpu().numpy()
        for i, obj in self.mesh_objects:
            if last[i] >= 0:
                obj._first_idx = int(last[i])


class DeviceColliders:
    """Device copy of a ColliderSet + the `d3d_colliders` struct pointing at it."""

    def __init__(self, cs, device):
        import torch
        to = lambda a: None if a is None else torch.from_numpy(a).to(device)  # noqa: E731
        mesh_last = None
        if cs.mesh_objects:
            mesh_last = torch.full((len(cs),), -1, dtype=torch.int32, device=device)
        self._init(to(cs.type), to(cs.pose), to(cs.param), to(cs.vert_off), to(cs.vert_len),
                   to(cs.verts), to(cs.margin), to(cs.graph_off), to(cs.graph), to(cs.mesh_start),
                   mesh_last)

    @classmethod
    def from_tensors(cls, type_, pose, param, vert_off=None, vert_len=None, verts=None, margin=None,
                     graph_off=None, graph=None, mesh_start=None, has_boxes=True):
        """Wrap device tensors that already live in HBM (no host round trip).  `has_boxes=False`:
        the caller knows that no collider is a Box, the d3d_prepare pass (box vertices) is skipped."""
        import torch
        self = cls.__new__(cls)
        n = type_.shape[0]
        dev = type_.device
        if vert_off is None:
            vert_off = torch.zeros(n, dtype=torch.int32, device=dev)
        if vert_len is None:
            vert_len = torch.zeros(n, dtype=torch.int32, device=dev)
        if verts is None:
            verts = torch.zeros((1, 3), dtype=torch.float64, device=dev)
        self._init(type_.to(torch.int32).contiguous(), pose.reshape(n, 4, 4).contiguous(),
                   param.reshape(n, 3).contiguous(), vert_off.contiguous(), vert_len.contiguous(),
                   verts.reshape(-1, 3).contiguous(), margin, graph_off, graph, mesh_start,
                   prepare=has_boxes)
        return self

    def _init(self, type_, pose, param, vert_off, vert_len, verts, margin, graph_off=None,
              graph=None, mesh_start=None, mesh_last=None, prepare=True):
        from . import _lib
        self.device = type_.device
        self.n = int(type_.shape[0])
        self.type, self.pose, self.param = type_, pose, param
        self.vert_off, self.vert_len, self.verts, self.margin = vert_off, vert_len, verts, margin
        self.struct = CColliders()
        self.struct.n = self.n
        self.struct.type = self.type.data_ptr()
        self.struct.pose = self.pose.data_ptr()
        self.struct.param = self.param.data_ptr()
        self.struct.vert_off = self.vert_off.data_ptr()
        self.struct.vert_len = self.vert_len.data_ptr()
        self.struct.verts = self.verts.data_ptr()
        self.struct.margin = None if self.margin is None else self.margin.data_ptr()
        if graph_off is None:
            graph = None
        self.graph_off, self.graph, self.mesh_start, self.mesh_last = graph_off, graph, mesh_start, mesh_last
        for name in ("graph_off", "graph", "mesh_start", "mesh_last"):
            t = getattr(self, name)
            if t is not None and t.dtype != torch_int32():
                raise TypeError("%s must be an int32 tensor" % name)
            setattr(self.struct, name, None if t is None else t.data_ptr())
        # box vertices are generated on the device (geometry.py:138-157)
        if prepare:
            _lib.prepare(self)


def concat_sets(sets):
    """Concatenate ColliderSets (vertex and adjacency pools are appended and re-based)."""
    n_v = n_g = 0
    vo, go = [], []
    any_graph = any(s.graph_off is not None for s in sets)
    any_margin = any(s.margin is not None for s in sets)
    for s in sets:
        vo.append(s.vert_off.astype(np.int64) + n_v)
        n_v += len(s.verts)
        if any_graph:
            if s.graph_off is None:
                go.append(np.full(len(s), -1, dtype=np.int64))
            else:
                go.append(np.where(s.graph_off >= 0, s.graph_off.astype(np.int64) + n_g, -1))
                n_g += len(s.graph)
    cat = lambda name: np.concatenate([getattr(s, name) for s in sets])  # noqa: E731
    graph = None
    if any_graph:
        # row pointers are relative to their record: the records are appended unchanged
        graph = np.concatenate([s.graph for s in sets if s.graph_off is not None])
    return ColliderSet(
        cat("type"), cat("pose"), cat("param"), np.concatenate(vo), cat("vert_len"), cat("verts"),
        np.concatenate([s.margin if s.margin is not None else np.zeros(len(s)) for s in sets])
        if any_margin else None,
        graph_off=np.concatenate(go) if any_graph else None, graph=graph,
        mesh_start=np.concatenate([s.mesh_start if s.mesh_start is not None
                                   else np.full(len(s), -1, dtype=np.int32) for s in sets])
        if any_graph else None)


def torch_int32():
    import torch
    return torch.int32


def _pose_from_center(center):
    T = np.eye(4)
    T[:3, 3] = center
    return T


def pack_colliders(colliders, track_mesh_state=False):
    """Pack a sequence of collider objects into a :class:`ColliderSet`.

    track_mesh_state: scalar-API mode - the set remembers its MeshGraph objects so that
    `commit_mesh_state` can carry their cached start vertex from call to call."""
    from . import colliders as C
    n = len(colliders)
    type_ = np.zeros(n, dtype=np.int32)
    pose = np.zeros((n, 4, 4))
    pose[:] = np.eye(4)
    param = np.zeros((n, 3))
    vert_off = np.zeros(n, dtype=np.int32)
    vert_len = np.zeros(n, dtype=np.int32)
    margin = np.zeros(n)
    any_margin = False
    chunks = []
    n_verts = 0
    graph_off = np.full(n, -1, dtype=np.int32)
    mesh_start = np.full(n, -1, dtype=np.int32)
    graph_chunks, graph_at, n_graph = [], {}, 0
    mesh_objects = []

    def add_vertices(i, V):
        nonlocal n_verts
        V = np.asarray(V, dtype=np.float64).reshape(-1, 3)
        vert_off[i] = n_verts
        vert_len[i] = len(V)
        chunks.append(V)
        n_verts += len(V)

    for i, c in enumerate(colliders):
        while isinstance(c, C.Margin):
            margin[i] += c.margin
            any_margin = True
            c = c.collider
        if isinstance(c, C.Box):
            type_[i] = BOX
            pose[i] = c.box2origin
            param[i] = c.size
            add_vertices(i, np.zeros((8, 3)))
        elif isinstance(c, C.ConvexHullVertices):
            type_[i] = HULL
            add_vertices(i, c.vertices)
        elif isinstance(c, C.MeshGraph):
            type_[i] = MESH
            pose[i] = c.mesh2origin
            add_vertices(i, c.vertices)
            g = c.graph_record()
            if id(g) not in graph_at:  # colliders that share a mesh share its record
                graph_at[id(g)] = n_graph
                graph_chunks.append(g)
                n_graph += len(g)
            graph_off[i] = graph_at[id(g)]
            if c._first_idx is not None:
                mesh_start[i] = c._first_idx
            mesh_objects.append((i, c))
        elif isinstance(c, C.Sphere):
            type_[i] = SPHERE
            pose[i] = _pose_from_center(c.c)
            param[i, 0] = c.radius
        elif isinstance(c, C.Capsule):
            type_[i] = CAPSULE
            pose[i] = c.capsule2origin
            param[i, :2] = (c.radius, c.height)
        elif isinstance(c, C.Ellipsoid):
            type_[i] = ELLIPSOID
            pose[i] = c.ellipsoid2origin
            param[i] = c.radii
        elif isinstance(c, C.Cylinder):
            type_[i] = CYLINDER
            pose[i] = c.cylinder2origin
            param[i, :2] = (c.radius, c.length)
        elif isinstance(c, C.Disk):
            type_[i] = DISK
            pose[i, :3, 3] = c.c
            pose[i, :3, 2] = c.normal
            param[i, 0] = c.radius
        elif isinstance(c, C.Ellipse):
            type_[i] = ELLIPSE
            pose[i, :3, 3] = c.c
            pose[i, :3, 0] = c.axes[0]
            pose[i, :3, 1] = c.axes[1]
            param[i, :2] = c.radii
        elif isinstance(c, C.Cone):
            type_[i] = CONE
            pose[i] = c.cone2origin
            param[i, :2] = (c.radius, c.height)
        else:
            raise TypeError("Unsupported collider type %r" % type(c))
    verts = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 3))
    cs = ColliderSet(type_, pose, param, vert_off, vert_len, verts,
                     margin if any_margin else None,
                     graph_off=graph_off if graph_chunks else None,
                     graph=np.concatenate(graph_chunks) if graph_chunks else None,
                     mesh_start=mesh_start if graph_chunks else None)
    if track_mesh_state and mesh_objects:
        cs.mesh_objects = mesh_objects
    return cs
