"""TEST INFRASTRUCTURE: ctypes wrapper of the CPU oracle (oracle/libd3d_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It consumes the same packed
`ColliderSet` host arrays as the CUDA library, so both sides see identical
inputs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_FLOAT = np.finfo(float).max
EPSILON = np.finfo(float).eps

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p


def build(force=False):
    so = os.path.join(_HERE, "libd3d_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.d3do_all_aabbs_overlap.restype = c_i64
        _LIB.d3do_tree_insert.restype = c_i64
        _LIB.d3do_tree_query.restype = c_i64
        _LIB.d3do_tree_vs_tree.restype = c_i64
        _LIB.d3do_max_threads.restype = c_int
    return _LIB


def _p(a):
    return c_ptr(a.ctypes.data)


def max_threads():
    """Host threads available to this process (torchrun exports OMP_NUM_THREADS=1, so
    the OpenMP default is not used: the thread count is passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def prepare(cs):
    """Fill box vertices in the host pool (geometry.py:138-157)."""
    s = cs.host_struct()
    lib().d3do_prepare(ctypes.byref(s), _p(cs.verts))
    cs.boxes_prepared = True
    return cs


def _ready(cs):
    if not cs.boxes_prepared:
        prepare(cs)
    return cs.host_struct()


def support(cs, idx, d, start=None, return_index=False):
    """support_function of collider idx.  MeshGraph: `start` = cached vertex of the
    object before the call (None: fresh object); return_index also returns the vertex the
    climb ended on (the object's new cache, mesh.py:85)."""
    s = _ready(cs)
    out = np.zeros(3)
    d = np.ascontiguousarray(d, dtype=np.float64)
    keep = []
    if start is not None:
        ms = np.full(len(cs), -1, dtype=np.int32)
        ms[idx] = start
        keep.append(ms)
        s.mesh_start = ms.ctypes.data
    last = np.full(len(cs), -1, dtype=np.int32)
    s.mesh_last = last.ctypes.data
    lib().d3do_support(ctypes.byref(s), c_i64(int(idx)), _p(d), _p(out))
    return (out, int(last[idx])) if return_index else out


def center(cs, idx):
    s = _ready(cs)
    out = np.zeros(3)
    lib().d3do_center(ctypes.byref(s), c_i64(int(idx)), _p(out))
    return out


def aabb(cs):
    s = _ready(cs)
    out = np.zeros((len(cs), 3, 2))
    lib().d3do_aabb(ctypes.byref(s), _p(out))
    return out


def _pairs(pairs):
    return np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)


def gjk_distance(cs, pairs, tolerance=1e-10, max_distance_squared=100000.0,
                 sanity_check=1e-8, n_threads=1):
    s = _ready(cs)
    pairs = _pairs(pairs)
    n = len(pairs)
    out = dict(dist=np.zeros(n), a=np.zeros((n, 3)), b=np.zeros((n, 3)), Y=np.zeros((n, 4, 3)),
               n_points=np.zeros(n, dtype=np.int32), iters=np.zeros(n, dtype=np.int32),
               status=np.zeros(n, dtype=np.int32))
    lib().d3do_gjk_distance(
        ctypes.byref(s), _p(pairs), c_i64(n), c_dbl(tolerance), c_dbl(max_distance_squared),
        c_dbl(sanity_check), _p(out["dist"]), _p(out["a"]), _p(out["b"]), _p(out["Y"]),
        _p(out["n_points"]), _p(out["iters"]), _p(out["status"]), c_int(n_threads))
    return out


def gjk_intersection(cs, pairs, tolerance=1e-10, n_threads=1):
    s = _ready(cs)
    pairs = _pairs(pairs)
    n = len(pairs)
    out = dict(hit=np.zeros(n, dtype=np.uint8), iters=np.zeros(n, dtype=np.int32),
               status=np.zeros(n, dtype=np.int32))
    lib().d3do_gjk_intersection(
        ctypes.byref(s), _p(pairs), c_i64(n), c_dbl(tolerance), _p(out["hit"]),
        _p(out["iters"]), _p(out["status"]), c_int(n_threads))
    return out


def epa(cs, pairs, Y, max_iter=64, max_loose_edges=32, max_faces=64, epsilon=1e-8,
        return_faces=False, n_threads=1):
    s = _ready(cs)
    pairs = _pairs(pairs)
    n = len(pairs)
    Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(n, 4, 3)
    out = dict(mtv=np.zeros((n, 3)), success=np.zeros(n, dtype=np.uint8),
               n_faces=np.zeros(n, dtype=np.int32), iters=np.zeros(n, dtype=np.int32),
               status=np.zeros(n, dtype=np.int32))
    faces = np.zeros((n, max_faces, 4, 3)) if return_faces else None
    lib().d3do_epa(
        ctypes.byref(s), _p(pairs), c_i64(n), _p(Y), c_int(max_iter), c_int(max_loose_edges),
        c_int(max_faces), c_dbl(epsilon), _p(out["mtv"]), _p(out["success"]),
        _p(out["n_faces"]), _p(out["iters"]), _p(out["status"]),
        _p(faces) if return_faces else None, c_int(n_threads))
    if return_faces:
        out["faces"] = faces
    return out


def mpr(cs, pairs, mpr_tolerance=0.0001, max_iterations=100, penetration=True, n_threads=1):
    s = _ready(cs)
    pairs = _pairs(pairs)
    n = len(pairs)
    out = dict(hit=np.zeros(n, dtype=np.uint8), depth=np.zeros(n), dir=np.zeros((n, 3)),
               pos=np.zeros((n, 3)), status=np.zeros(n, dtype=np.int32))
    lib().d3do_mpr(
        ctypes.byref(s), _p(pairs), c_i64(n), c_dbl(mpr_tolerance), c_int(max_iterations),
        c_int(1 if penetration else 0), _p(out["hit"]), _p(out["depth"]), _p(out["dir"]),
        _p(out["pos"]), _p(out["status"]), c_int(n_threads))
    return out


def all_aabbs_overlap(aabbs1, aabbs2):
    """Brute force (aabb_tree.py:465-500); returns pairs int32[k,2] in row-major order."""
    a1 = np.ascontiguousarray(aabbs1, dtype=np.float64).reshape(-1, 3, 2)
    a2 = np.ascontiguousarray(aabbs2, dtype=np.float64).reshape(-1, 3, 2)
    n = lib().d3do_all_aabbs_overlap(_p(a1), c_i64(len(a1)), _p(a2), c_i64(len(a2)), None, c_i64(0))
    pairs = np.zeros((max(n, 1), 2), dtype=np.int32)
    lib().d3do_all_aabbs_overlap(_p(a1), c_i64(len(a1)), _p(a2), c_i64(len(a2)), _p(pairs), c_i64(n))
    return pairs[:n]


class Tree:
    """Incremental AABB tree of the reference (aabb_tree.py:15-341)."""

    def __init__(self):
        self.root = -1
        self.filled_len = 0
        self.nodes = np.empty((0, 4), dtype=np.int64)
        self.aabbs = np.empty((0, 3, 2))

    def insert_aabbs(self, aabbs):
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float64).reshape(-1, 3, 2)
        n = len(aabbs)
        if n == 0:
            return
        old = self.filled_len
        self.filled_len += n
        new_nodes = np.full((2 * (self.filled_len - len(self.nodes)), 4), -1, dtype=np.int64)
        self.nodes = np.ascontiguousarray(np.append(self.nodes, new_nodes, axis=0))
        self.aabbs = np.append(self.aabbs, aabbs, axis=0)
        self.aabbs = np.ascontiguousarray(np.append(
            self.aabbs, np.zeros((len(self.nodes) - len(self.aabbs), 3, 2)), axis=0))
        order = np.arange(old, self.filled_len, dtype=np.int64)
        fl = c_i64(self.filled_len)
        self.root = int(lib().d3do_tree_insert(
            c_i64(self.root), _p(self.nodes), _p(self.aabbs), ctypes.byref(fl), _p(order), c_i64(n)))
        self.filled_len = int(fl.value)
        self.nodes = np.ascontiguousarray(self.nodes[:self.filled_len])
        self.aabbs = np.ascontiguousarray(self.aabbs[:self.filled_len])

    def query(self, query_aabbs, n_threads=1):
        """Per-box query_overlap (aabb_tree.py:381-403); returns (leaf, query) int32[k,2]."""
        q = np.ascontiguousarray(query_aabbs, dtype=np.float64).reshape(-1, 3, 2)
        n = lib().d3do_tree_query(c_i64(self.root), _p(self.nodes), _p(self.aabbs), _p(q),
                                  c_i64(len(q)), None, c_i64(0), c_int(n_threads))
        pairs = np.zeros((max(n, 1), 2), dtype=np.int32)
        lib().d3do_tree_query(c_i64(self.root), _p(self.nodes), _p(self.aabbs), _p(q),
                              c_i64(len(q)), _p(pairs), c_i64(n), c_int(n_threads))
        return pairs[:n]

    def overlaps_aabb_tree(self, other):
        """query_overlap_of_other_tree (aabb_tree.py:344-378); pairs (idx_self, idx_other)."""
        n = lib().d3do_tree_vs_tree(c_i64(self.root), _p(self.nodes), _p(self.aabbs),
                                    c_i64(other.root), _p(other.nodes), _p(other.aabbs),
                                    None, c_i64(0))
        pairs = np.zeros((max(n, 1), 2), dtype=np.int32)
        lib().d3do_tree_vs_tree(c_i64(self.root), _p(self.nodes), _p(self.aabbs),
                                c_i64(other.root), _p(other.nodes), _p(other.aabbs),
                                _p(pairs), c_i64(n))
        return pairs[:n]


def norm(v):
    """BLAS dnrm2 (x87) of each row of v[n,3]."""
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(len(v))
    lib().d3do_norm(_p(v), c_i64(len(v)), _p(out))
    return out


def fk(kin, q, n_threads=1):
    """Poses [B,K,4,4] of a flattened URDF chain (`UrdfTransformManager.compile_kinematics`)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    n_joints = len(kin["joint_names"])
    n_frames = len(kin["chain_off"]) - 1
    q = q.reshape(-1, n_joints)
    out = np.zeros((len(q), n_frames, 4, 4))
    arrs = {k: np.ascontiguousarray(kin[k]) for k in ("joint_axis", "joint_limits", "joint_type",
                                                      "chain_off", "chain_fixed", "chain_joint")}
    lib().d3do_fk(c_int(n_frames), c_int(n_joints), _p(arrs["joint_axis"]), _p(arrs["joint_limits"]),
                  _p(arrs["joint_type"]), _p(arrs["chain_off"]), _p(arrs["chain_fixed"]),
                  _p(arrs["chain_joint"]), _p(q), c_i64(len(q)), _p(out), c_int(n_threads))
    return out


def self_collision_masks(template, kin, pattern, q, n_threads=1):
    """Reference semantics of self_collision.detect (self_collision.py:5-36) for many joint
    configurations: FK, AABBs, white-list filtered candidates, gjk_intersection.
    `template`: ColliderSet of the K colliders (shape parameters), `pattern` int32[C,2]."""
    from distance3d_b200.pack import ColliderSet
    poses = fk(kin, q, n_threads)
    B, K = poses.shape[:2]
    z = np.zeros(B * K, dtype=np.int32)
    cs = ColliderSet(np.tile(template.type, B), poses.reshape(-1, 4, 4), np.tile(template.param, (B, 1)),
                     z, z, np.zeros((0, 3)))
    boxes = aabb(cs).reshape(B, K, 3, 2)
    a, b = boxes[:, pattern[:, 0]], boxes[:, pattern[:, 1]]
    ov = np.all((a[..., 0] <= b[..., 1]) & (a[..., 1] >= b[..., 0]), axis=-1)   # [B, C]
    bi, ci = np.nonzero(ov)
    pairs = np.stack((bi * K + pattern[ci, 0], bi * K + pattern[ci, 1]), axis=1).astype(np.int32)
    hit = gjk_intersection(cs, pairs, n_threads=n_threads)["hit"].astype(bool)
    mask = np.zeros(B * K, dtype=np.uint8)
    mask[pairs[hit, 0]] = 1
    mask[pairs[hit, 1]] = 1
    return mask.reshape(B, K), len(pairs)


def self_collision_masks_ordered(template, kin, whitelist, q, n_threads=1):
    """self_collision.detect (self_collision.py:22-36) replayed literally for every joint
    configuration, candidate order included: the BVH is rebuilt by inserting the colliders
    one at a time into the reference's incremental tree (broad_phase.py:144-151), the
    candidates of a frame are the tree's overlaps in query order (aabb_tree.py:381-403)
    minus its white-list, the first intersecting candidate flags both frames.
    `whitelist` uint8[K,K]: whitelist[i, j] = frame j is white-listed for frame i."""
    from distance3d_b200.pack import ColliderSet
    poses = fk(kin, q, n_threads)
    B, K = poses.shape[:2]
    z = np.zeros(K, dtype=np.int32)
    masks = np.zeros((B, K), dtype=np.uint8)
    for b in range(B):
        cs = ColliderSet(template.type, poses[b], template.param, z, z, np.zeros((0, 3)))
        boxes = aabb(cs)
        tree = Tree()
        leaf_of = []
        for k in range(K):
            leaf_of.append(tree.filled_len)
            tree.insert_aabbs(boxes[k:k + 1])
        collider_of = {leaf: k for k, leaf in enumerate(leaf_of)}
        contacts = {}
        for f in range(K):
            if f in contacts:
                continue
            cand = [collider_of[int(leaf)] for leaf, _ in tree.query(boxes[f:f + 1])]
            contacts[f] = False
            for g2 in cand:
                if whitelist[f, g2]:
                    continue
                if f == g2 or gjk_intersection(cs, [[f, g2]])["hit"][0]:
                    contacts[f] = True
                    contacts[g2] = True
                    break
        masks[b] = [contacts[f] for f in range(K)]
    return masks


def tetra_aabbs(points):
    """tetrahedral_mesh_aabbs (hydroelastic_contact/_mesh_processing.py:4-20)."""
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4, 3)
    out = np.zeros((len(points), 3, 2))
    lib().d3do_tetra_aabbs(_p(points), c_i64(len(points)), _p(out))
    return out


def barycentric_transforms(points):
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 4, 3)
    X = np.zeros((len(points), 4, 4))
    for k in range(len(points)):
        lib().d3do_barycentric_transform(_p(points[k]), _p(X[k]))
    return X


def tetra_pairs(pairs, points1, eps1, points2, eps2, X1=None, X2=None, youngs_modulus1=1.0,
                youngs_modulus2=1.0, max_vertices=12, n_threads=1):
    """intersect_tetrahedron_pairs (hydroelastic_contact/_tetrahedron_intersection.py:7-140)."""
    pairs = _pairs(pairs)
    n = len(pairs)
    f = lambda a, shape: np.ascontiguousarray(a, dtype=np.float64).reshape(shape)  # noqa: E731
    points1, points2 = f(points1, (-1, 4, 3)), f(points2, (-1, 4, 3))
    eps1, eps2 = f(eps1, (-1, 4)), f(eps2, (-1, 4))
    X1 = None if X1 is None else f(X1, (-1, 4, 4))
    X2 = None if X2 is None else f(X2, (-1, 4, 4))
    out = dict(hit=np.zeros(n, dtype=np.uint8), plane=np.zeros((n, 4)), n_vertices=np.zeros(n, dtype=np.int32),
               polygon=np.zeros((n, max_vertices, 3)), status=np.zeros(n, dtype=np.int32))
    lib().d3do_tetra_pairs(_p(pairs), c_i64(n), _p(points1), _p(eps1), None if X1 is None else _p(X1),
                           _p(points2), _p(eps2), None if X2 is None else _p(X2),
                           c_dbl(youngs_modulus1), c_dbl(youngs_modulus2), c_int(max_vertices),
                           _p(out["hit"]), _p(out["plane"]), _p(out["n_vertices"]), _p(out["polygon"]),
                           _p(out["status"]), c_int(n_threads))
    return out


def gjk_intersection_libccd(cs, pairs, max_iterations=100, n_threads=1):
    """gjk_intersection_libccd (gjk/_gjk_libccd.py:14-91)."""
    s = _ready(cs)
    pairs = _pairs(pairs)
    n = len(pairs)
    out = dict(hit=np.zeros(n, dtype=np.uint8), iters=np.zeros(n, dtype=np.int32))
    lib().d3do_gjk_intersection_libccd(ctypes.byref(s), _p(pairs), c_i64(n), c_int(max_iterations),
                                       _p(out["hit"]), _p(out["iters"]), c_int(n_threads))
    return out
