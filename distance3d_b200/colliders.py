"""Collider classes with the reference's constructor signatures and attributes.

Drop-in for distance3d/colliders.py:17-659.  The objects are thin parameter
holders; every geometric query (`support_function`, `aabb`, and the narrow /
broad phase built on them) is answered by the CUDA library on a batch of one.
Use :func:`distance3d_b200.pack.pack_colliders` and the ``*_batch`` entry
points for throughput.
"""
import numpy as np

from . import pack as _pack


class ConvexCollider:
    """Base class (reference: colliders.py:17-106)."""

    def __init__(self, artist=None):
        self.artist_ = artist

    # -- packed single-collider view -------------------------------------
    def _set(self):
        # attributes may be mutated freely (as in the reference), so the
        # one-collider record is re-packed on every scalar query
        return _pack.pack_colliders([self], track_mesh_state=True)

    def _dirty(self):
        pass

    def make_artist(self, c=None):
        raise NotImplementedError(
            "visualisation is out of scope of distance3d_b200 (SURVEY.md section 2)")

    def support_function(self, search_direction):
        """Extreme point along `search_direction` (computed on the GPU)."""
        from . import _lib
        d = np.ascontiguousarray(search_direction, dtype=np.float64).reshape(1, 3)
        cs = self._set()
        out = _lib.support(cs, np.zeros(1, dtype=np.int32), d)[0]
        cs.commit_mesh_state()  # MeshGraph (also inside a Margin) caches its last vertex
        return out

    def aabb(self):
        """Axis-aligned bounding box, shape (3, 2) (computed on the GPU)."""
        from . import _lib
        return _lib.aabb(self._set())[0]

    def first_vertex(self):
        raise NotImplementedError

    def center(self):
        raise NotImplementedError

    def update_pose(self, pose):
        raise NotImplementedError

    def collider2origin(self):
        raise NotImplementedError


class ConvexHullVertices(ConvexCollider):
    """Convex hull of world-frame vertices (reference: colliders.py:109-144)."""

    def __init__(self, vertices, artist=None):
        super().__init__(artist)
        self.vertices = vertices

    def first_vertex(self):
        return self.vertices[0]

    def center(self):
        return np.mean(self.vertices, axis=0)

    def update_pose(self, pose):
        raise NotImplementedError("update_pose is not implemented!")

    def collider2origin(self):
        return np.eye(4)


class Box(ConvexHullVertices):
    """Oriented box (reference: colliders.py:147-184).

    The reference turns a box into 8 world-frame vertices on the host; here the
    vertices are produced on the device by `d3d_prepare` when the collider is
    packed.  `vertices` is available lazily for API compatibility.
    """

    def __init__(self, box2origin, size, artist=None):
        ConvexCollider.__init__(self, artist)
        self.box2origin = box2origin
        self.size = size

    @property
    def vertices(self):
        from . import _lib
        return _lib.box_vertices(self._set())[0]

    @vertices.setter
    def vertices(self, value):  # pragma: no cover - kept for attribute parity
        raise AttributeError("Box vertices are derived from pose and size")

    def first_vertex(self):
        return self.vertices[0]

    def center(self):
        return self.box2origin[:3, 3]

    def update_pose(self, pose):
        self.box2origin = pose
        self._dirty()

    def collider2origin(self):
        return self.box2origin


class MeshGraph(ConvexCollider):
    """Convex mesh in its own frame (reference: colliders.py:187-240).

    The support map climbs the triangle graph (mesh.py:12-139): six axis-extreme shortcut
    vertices, then neighbour rounds with the 10*EPSILON plateau rule, starting from the
    vertex the previous support call ended on.  The adjacency record is built once here
    (mesh.build_mesh_graph) and uploaded with the collider; like the reference object,
    this one remembers its last vertex across scalar calls (`_first_idx`, mesh.py:85).
    In batched calls every pair starts from the packed start vertex.
    """

    def __init__(self, mesh2origin, vertices, triangles, artist=None):
        super().__init__(artist)
        self.mesh2origin = mesh2origin
        self.vertices = vertices
        self.triangles = triangles
        self._graph = None
        self._first_idx = None  # None = min(triangles), the fresh object's start

    def graph_record(self):
        if self._graph is None:
            from .mesh import build_mesh_graph
            self._graph = build_mesh_graph(self.vertices, self.triangles)
        return self._graph

    def first_vertex(self):
        return self.mesh2origin[:3, 3] + np.dot(self.mesh2origin[:3, :3], self.vertices[0])

    def center(self):
        return self.mesh2origin[:3, 3] + np.dot(
            self.mesh2origin[:3, :3], np.mean(self.vertices, axis=0))

    def update_pose(self, mesh2origin):
        self.mesh2origin = mesh2origin
        self._dirty()

    def collider2origin(self):
        return self.mesh2origin


class Sphere(ConvexCollider):
    """Sphere (reference: colliders.py:243-287)."""

    def __init__(self, center, radius, artist=None):
        super().__init__(artist)
        self.c = center
        self.radius = radius

    def center(self):
        return self.c

    def first_vertex(self):
        return self.c + np.array([0, 0, self.radius], dtype=float)

    def update_pose(self, pose):
        self.c = pose[:3, 3]
        self._dirty()

    def collider2origin(self):
        sphere2origin = np.eye(4)
        sphere2origin[:3, 3] = self.c
        return sphere2origin


class Capsule(ConvexCollider):
    """Capsule along local z (reference: colliders.py:290-340)."""

    def __init__(self, capsule2origin, radius, height, artist=None):
        super().__init__(artist)
        self.capsule2origin = capsule2origin
        self.radius = radius
        self.height = height

    def center(self):
        return self.capsule2origin[:3, 3]

    def first_vertex(self):
        return self.capsule2origin[:3, 3] - (
            self.radius + 0.5 * self.height) * self.capsule2origin[:3, 2]

    def update_pose(self, pose):
        self.capsule2origin = pose
        self._dirty()

    def collider2origin(self):
        return self.capsule2origin


class Ellipsoid(ConvexCollider):
    """Ellipsoid (reference: colliders.py:343-387)."""

    def __init__(self, ellipsoid2origin, radii, artist=None):
        super().__init__(artist)
        self.ellipsoid2origin = ellipsoid2origin
        self.radii = radii

    def center(self):
        return self.ellipsoid2origin[:3, 3]

    def first_vertex(self):
        return self.ellipsoid2origin[:3, 3] + self.radii[2] * self.ellipsoid2origin[:3, 2]

    def update_pose(self, pose):
        self.ellipsoid2origin = pose
        self._dirty()

    def collider2origin(self):
        return self.ellipsoid2origin


class Cylinder(ConvexCollider):
    """Cylinder along local z (reference: colliders.py:390-440)."""

    def __init__(self, cylinder2origin, radius, length, artist=None):
        super().__init__(artist)
        self.cylinder2origin = cylinder2origin
        self.radius = radius
        self.length = length

    def center(self):
        return self.cylinder2origin[:3, 3]

    def first_vertex(self):
        return self.cylinder2origin[:3, 3] + 0.5 * self.length * self.cylinder2origin[:3, 2]

    def update_pose(self, pose):
        self.cylinder2origin = pose
        self._dirty()

    def collider2origin(self):
        return self.cylinder2origin


class Disk(ConvexCollider):
    """Disk (reference: colliders.py:443-497)."""

    def __init__(self, center, radius, normal, artist=None):
        super().__init__(artist)
        self.c = center
        self.radius = radius
        self.normal = normal

    def center(self):
        return self.c

    def update_pose(self, pose):
        self.c = pose[:3, 3]
        self.normal = pose[:3, 2]
        self._dirty()

    def collider2origin(self):
        from ._transforms import plane_basis_from_normal
        x, y = plane_basis_from_normal(self.normal)
        disk2origin = np.eye(4)
        disk2origin[:3, :3] = np.column_stack((x, y, self.normal))
        disk2origin[:3, 3] = self.c
        return disk2origin

    def first_vertex(self):
        from ._transforms import plane_basis_from_normal
        x, _ = plane_basis_from_normal(self.normal)
        return self.c + self.radius * x


class Ellipse(ConvexCollider):
    """Ellipse (reference: colliders.py:500-551)."""

    def __init__(self, center, axes, radii, artist=None):
        super().__init__(artist)
        self.c = center
        self.axes = axes
        self.radii = radii

    def center(self):
        return self.c

    def first_vertex(self):
        return self.c + self.axes[0] * self.radii[0]

    def update_pose(self, pose):
        self.c = pose[:3, 3]
        self.axes = pose[:3, :2].T
        self._dirty()

    def collider2origin(self):
        ellipse2origin = np.eye(4)
        ellipse2origin[:3, :2] = np.asarray(self.axes).T
        ellipse2origin[:3, 2] = np.cross(self.axes[0], self.axes[1])
        ellipse2origin[:3, 3] = self.c
        return ellipse2origin


class Cone(ConvexCollider):
    """Cone with base at the local origin, apex at +z*height (colliders.py:554-603)."""

    def __init__(self, cone2origin, radius, height, artist=None):
        super().__init__(artist)
        self.cone2origin = cone2origin
        self.radius = radius
        self.height = height

    def center(self):
        return self.cone2origin[:3, 3] + 0.5 * self.height * self.cone2origin[:3, 2]

    def first_vertex(self):
        return self.cone2origin[:3, 3] + self.height * self.cone2origin[:3, 2]

    def update_pose(self, pose):
        self.cone2origin = pose
        self._dirty()

    def collider2origin(self):
        return self.cone2origin


class Margin(ConvexCollider):
    """Margin around another collider (reference: colliders.py:606-646)."""

    def __init__(self, collider, margin):
        super().__init__(collider.artist_)
        self.collider = collider
        self.margin = margin

    def first_vertex(self):
        return self.collider.first_vertex()

    def center(self):
        return self.collider.center()

    def update_pose(self, pose):
        self.collider.update_pose(pose)
        self._dirty()

    def collider2origin(self):
        return self.collider.collider2origin()


COLLIDERS = {
    "sphere": Sphere,
    "ellipsoid": Ellipsoid,
    "capsule": Capsule,
    "disk": Disk,
    "ellipse": Ellipse,
    "cone": Cone,
    "cylinder": Cylinder,
    "box": Box,
    "mesh": MeshGraph,
}
