"""Mesh input (reference: distance3d/io.py:5-46 `load_mesh`).

The reference delegates to Open3D or trimesh; neither is a dependency here.  The two formats
its data and examples use for collision geometry are read directly with numpy: STL (binary
and ASCII) and Wavefront OBJ.  A `MeshGraph` climbs the triangle graph (mesh.py:12-139), which
only works on a connected mesh, so vertices that occur several times in the file - every STL
facet repeats its corners - are merged, keeping the order of first occurrence.
"""
import os
import struct

import numpy as np


def _merge_vertices(corners):
    """corners[n_triangles * 3, 3] -> (unique vertices in order of first occurrence, triangles)."""
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 3)
    if len(corners) == 0:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=int)
    _, first, inverse = np.unique(corners, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    order = np.argsort(first, kind="stable")          # unique id -> rank by first occurrence
    rank = np.empty(len(order), dtype=int)
    rank[order] = np.arange(len(order))
    vertices = corners[first[order]]
    triangles = rank[inverse].reshape(-1, 3)
    return vertices, triangles


def _load_stl_binary(data):
    n = struct.unpack_from("<I", data, 80)[0]
    rec = np.dtype([("normal", "<f4", 3), ("corners", "<f4", (3, 3)), ("attr", "<u2")])
    if len(data) < 84 + n * rec.itemsize:
        raise OSError("truncated binary STL: %d facets announced, %d bytes" % (n, len(data)))
    facets = np.frombuffer(data, dtype=rec, count=n, offset=84)
    return _merge_vertices(facets["corners"].astype(np.float64))


def _load_stl_ascii(text):
    corners = []
    for line in text.splitlines():
        parts = line.split()
        if len(parts) == 4 and parts[0].lower() == "vertex":
            corners.append([float(parts[1]), float(parts[2]), float(parts[3])])
    if len(corners) % 3 != 0:
        raise OSError("ASCII STL with %d vertices (not a multiple of 3)" % len(corners))
    return _merge_vertices(np.array(corners, dtype=np.float64).reshape(-1, 3))


def _load_stl(filename):
    with open(filename, "rb") as f:
        data = f.read()
    if len(data) >= 84:
        n = struct.unpack_from("<I", data, 80)[0]
        if len(data) == 84 + 50 * n:      # the size test is the reliable one: binary files may start with "solid"
            return _load_stl_binary(data)
    head = data[:512].lstrip().lower()
    if head.startswith(b"solid"):
        return _load_stl_ascii(data.decode("ascii", errors="replace"))
    if len(data) >= 84:
        return _load_stl_binary(data)
    raise OSError("not an STL file: %s" % filename)


def _load_obj(filename):
    vertices, triangles = [], []
    with open(filename, "r", errors="replace") as f:
        for line in f:
            parts = line.split()
            if not parts:
                continue
            if parts[0] == "v":
                vertices.append([float(x) for x in parts[1:4]])
            elif parts[0] == "f":
                idx = [int(p.split("/")[0]) for p in parts[1:]]
                idx = [i - 1 if i > 0 else len(vertices) + i for i in idx]
                for k in range(1, len(idx) - 1):   # fan triangulation of polygons
                    triangles.append([idx[0], idx[k], idx[k + 1]])
    return (np.array(vertices, dtype=np.float64).reshape(-1, 3),
            np.array(triangles, dtype=int).reshape(-1, 3))


def load_mesh(filename, scale=1.0):
    """Load a triangle mesh (io.py:5-46).

    Parameters
    ----------
    filename : str
        Path to an ``.stl`` (binary or ASCII) or ``.obj`` file.

    scale : float or array-like, shape (3,), optional (default: 1)
        Scale of vertex coordinates (URDF ``<mesh scale="sx sy sz">``).

    Returns
    -------
    vertices : array, shape (n_vertices, 3)
    triangles : array, shape (n_triangles, 3)
        Indices of vertices that form triangles of the mesh.
    """
    ext = os.path.splitext(filename)[1].lower()
    if ext == ".stl":
        vertices, triangles = _load_stl(filename)
    elif ext == ".obj":
        vertices, triangles = _load_obj(filename)
    else:
        raise OSError("unsupported mesh format '%s' (%s); supported: .stl, .obj" % (ext, filename))
    vertices = vertices * np.asarray(scale, dtype=np.float64)
    return vertices, triangles


def save_stl(filename, vertices, triangles, binary=True):
    """Write a triangle mesh as STL (test fixtures, round trips)."""
    V = np.asarray(vertices, dtype=np.float64)
    T = np.asarray(triangles, dtype=int).reshape(-1, 3)
    corners = V[T]                                                    # [n, 3, 3]
    normals = np.cross(corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0])
    length = np.linalg.norm(normals, axis=1)
    normals = normals / np.where(length > 0.0, length, 1.0)[:, None]
    if binary:
        rec = np.dtype([("normal", "<f4", 3), ("corners", "<f4", (3, 3)), ("attr", "<u2")])
        facets = np.zeros(len(T), dtype=rec)
        facets["normal"] = normals
        facets["corners"] = corners
        with open(filename, "wb") as f:
            f.write(b"distance3d_b200".ljust(80, b" "))
            f.write(struct.pack("<I", len(T)))
            f.write(facets.tobytes())
    else:
        with open(filename, "w") as f:
            f.write("solid mesh\n")
            for n, c in zip(normals, corners):
                f.write("facet normal %r %r %r\n outer loop\n" % tuple(float(x) for x in n))
                for v in c:
                    f.write("  vertex %r %r %r\n" % tuple(float(x) for x in v))
                f.write(" endloop\nendfacet\n")
            f.write("endsolid mesh\n")
