"""Convex-mesh helpers (reference: distance3d/mesh.py:142-221).

`MeshSupportFunction` is the brute-force support map; the reference's hill-climbing
variant (mesh.py:12-139) gives the same point except on 10*EPSILON plateaus and is what
`colliders.MeshGraph` replaces by the device arg-max."""
import numpy as np

from .utils import HALF_PI


class MeshSupportFunction:
    """Support function of a convex mesh: arg-max over all vertices (mesh.py:142-191)."""

    def __init__(self, mesh2origin, vertices, triangles):
        self.mesh2origin = mesh2origin
        self.vertices = vertices
        self.triangles = triangles
        self.first_idx = 0

    def update_pose(self, mesh2origin):
        self.mesh2origin = mesh2origin

    def __call__(self, search_direction):
        """Returns (index of the support vertex, support point in the origin frame)."""
        from .colliders import MeshGraph
        point = MeshGraph(self.mesh2origin, self.vertices, self.triangles).support_function(
            search_direction)
        local = np.dot(np.asarray(self.mesh2origin)[:3, :3].T, point - np.asarray(self.mesh2origin)[:3, 3])
        idx = int(np.argmin(np.linalg.norm(np.asarray(self.vertices) - local, axis=1)))
        return idx, point


MeshHillClimbingSupportFunction = MeshSupportFunction


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
