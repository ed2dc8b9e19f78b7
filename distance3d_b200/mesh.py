"""Convex-mesh helpers (reference: distance3d/mesh.py:12-221).

`build_mesh_graph` turns a triangle list into the adjacency record the kernels climb
(include/d3d_types.h); `MeshHillClimbingSupportFunction` / `MeshSupportFunction` are the
reference's two support-map objects, both answered by the CUDA library."""
import numpy as np

from .utils import HALF_PI


PROJECTION_LENGTH_EPSILON = 10.0 * np.finfo(float).eps  # mesh.py:9


def build_mesh_graph(vertices, triangles):
    """int32 adjacency record of a triangle mesh (layout: include/d3d_types.h).

    The climb visits the neighbours of a vertex in the order the reference iterates its
    per-vertex Python `set` (mesh.py:31-41, 49-52), and that order decides which of several
    improving neighbours is taken first.  A set's iteration order is a function of the
    hashes and of the insertion sequence only, so the neighbour sets are filled here by the
    same sequence of insertions (per triangle: j, k into i; i, k into j; i, j into k) and
    read back with plain iteration.
    """
    V = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    tri = np.asarray(triangles).reshape(-1, 3)
    nv = len(V)
    if len(tri) and (tri.min() < 0 or tri.max() >= nv):
        raise ValueError("triangle index out of range")
    rings = [set() for _ in range(nv)]
    for a, b, c in tri.tolist():
        rings[a].add(b), rings[a].add(c)
        rings[b].add(a), rings[b].add(c)
        rings[c].add(a), rings[c].add(b)
    counts = np.fromiter((len(r) for r in rings), dtype=np.int64, count=nv)
    head = 7 + nv + 1
    g = np.empty(head + int(counts.sum()), dtype=np.int32)
    g[0] = int(tri.min()) if len(tri) else 0                          # mesh.py:29
    g[1:4] = np.argmax(V, axis=0)                                     # mesh.py:44-47
    g[4:7] = np.argmin(V, axis=0)
    g[7] = head
    g[8:head] = head + np.cumsum(counts)
    pos = head
    for r in rings:
        g[pos:pos + len(r)] = list(r)
        pos += len(r)
    return g


class _DeviceMeshSupport:
    def __init__(self, mesh2origin, vertices, triangles):
        self.mesh2origin = mesh2origin
        self.vertices = vertices
        self.triangles = triangles

    def update_pose(self, mesh2origin):
        self.mesh2origin = mesh2origin

    def _index_of(self, point):
        T = np.asarray(self.mesh2origin)
        local = np.dot(T[:3, :3].T, point - T[:3, 3])
        return int(np.argmin(np.linalg.norm(np.asarray(self.vertices) - local, axis=1)))


class MeshHillClimbingSupportFunction(_DeviceMeshSupport):
    """Hill-climbing support map with vertex caching (mesh.py:12-87)."""

    def __init__(self, mesh2origin, vertices, triangles):
        super().__init__(mesh2origin, vertices, triangles)
        from .colliders import MeshGraph
        self._mesh = MeshGraph(mesh2origin, vertices, triangles)
        self.first_idx = int(np.min(triangles))
        self.shortcut_connections = self._mesh.graph_record()[1:7].astype(np.int64)

    def __call__(self, search_direction):
        """Returns (index of the support vertex, support point in the origin frame)."""
        self._mesh.mesh2origin = self.mesh2origin
        self._mesh._first_idx = self.first_idx
        point = self._mesh.support_function(search_direction)
        self.first_idx = self._mesh._first_idx
        return self.first_idx, point


class MeshSupportFunction(_DeviceMeshSupport):
    """Arg-max over all vertices (mesh.py:142-191)."""

    def __init__(self, mesh2origin, vertices, triangles):
        super().__init__(mesh2origin, vertices, triangles)
        self.first_idx = 0

    def __call__(self, search_direction):
        """Returns (index of the support vertex, support point in the origin frame)."""
        from . import _lib, pack
        cs = pack.ColliderSet([pack.MESH], [self.mesh2origin], np.zeros((1, 3)), [0],
                              [len(self.vertices)], self.vertices)  # no graph: arg-max
        d = np.ascontiguousarray(search_direction, dtype=np.float64).reshape(1, 3)
        point = _lib.support(cs, np.zeros(1, dtype=np.int32), d)[0]
        return self._index_of(point), point


def hill_climb_mesh_extreme(search_direction, start_idx, vertices, connections,
                            shortcut_connections):
    raise NotImplementedError(
        "the climb runs inside the CUDA support map (csrc/d3d_support.cuh hill_climb); "
        "use MeshHillClimbingSupportFunction")


def make_convex_mesh(vertices):
    """Triangles of the convex hull with outward normals (mesh.py:194-221); input generation."""
    from scipy.spatial import ConvexHull
    vertices = vertices - np.mean(vertices, axis=0)
    triangles = ConvexHull(vertices).simplices
    faces = vertices[triangles]
    normals = np.cross(faces[:, 2] - faces[:, 0], faces[:, 1] - faces[:, 0])
    centers = np.mean(faces, axis=1)
    cosang = np.sum(normals * centers, axis=1) / (
        np.linalg.norm(normals, axis=1) * np.linalg.norm(centers, axis=1))
    angles = np.arccos(np.clip(cosang, -1.0, 1.0))
    flip = np.where(angles < HALF_PI)[0]
    triangles[flip] = triangles[flip, ::-1]
    return triangles
