"""Random shapes: scalar generators with the reference's signatures
(distance3d/random.py:200-494) and a vectorised generator that emits a packed
:class:`ColliderSet` directly (millions of shapes without Python loops).

Poses come from :func:`_transforms.random_transform` (exponential coordinates
~ N(0, I6)), the restatement of ``pytransform3d.transformations.random_transform``
that the reference calls (random.py:222,260,298,340,394,450).  Only the
distribution matters: oracle and CUDA path consume the same arrays.
"""
import numpy as np

from . import pack as _pack
from ._transforms import norm_vector, perpendicular_to_vector, random_transform


def randn_point(random_state, scale=1.0):
    return scale * random_state.randn(3)


def randn_direction(random_state):
    return norm_vector(random_state.randn(3))


def randn_rectangle(random_state, center_scale=1.0, length_scale=1.0):
    center = center_scale * randn_point(random_state)
    axis1 = randn_direction(random_state)
    axis2 = norm_vector(perpendicular_to_vector(axis1))
    lengths = (1.0 - random_state.rand(2)) * length_scale
    return center, np.vstack((axis1, axis2)), lengths


def rand_circle(random_state, radius_scale=1.0):
    center = random_state.randn(3)
    radius = (1.0 - random_state.rand()) * radius_scale
    normal = norm_vector(random_state.randn(3))
    return center, radius, normal


def rand_box(random_state, center_scale=1.0, size_scale=1.0):
    box2origin = random_transform(random_state)
    box2origin[:3, 3] *= center_scale
    size = (1.0 - random_state.rand(3)) * size_scale
    return box2origin, size


def rand_capsule(random_state, center_scale=1.0, radius_scale=1.0, height_scale=1.0):
    capsule2origin = random_transform(random_state)
    capsule2origin[:3, 3] *= center_scale
    radius = (1.0 - random_state.rand()) * radius_scale
    height = (1.0 - random_state.rand()) * height_scale
    return capsule2origin, radius, height


def rand_ellipsoid(random_state, center_scale=1.0, min_radius=0.0, radius_scale=1.0):
    ellipsoid2origin = random_transform(random_state)
    ellipsoid2origin[:3, 3] *= center_scale
    radii = min_radius + (1.0 - random_state.rand(3)) * radius_scale
    return ellipsoid2origin, radii


def rand_cylinder(random_state, center_scale=1.0, min_radius=0.0, min_length=0.0,
                  radius_scale=1.0, length_scale=1.0):
    cylinder2origin = random_transform(random_state)
    cylinder2origin[:3, 3] *= center_scale
    radius = min_radius + (1.0 - random_state.rand()) * radius_scale
    length = min_length + (1.0 - random_state.rand()) * length_scale
    return cylinder2origin, radius, length


def rand_sphere(random_state, center_scale=1.0, radius_scale=1.0):
    center = random_state.randn(3) * center_scale
    radius = (1.0 - random_state.rand()) * radius_scale
    return center, radius


def rand_cone(random_state, center_scale=1.0, min_radius=0.0, min_height=0.0,
              radius_scale=1.0, height_scale=1.0):
    cone2origin = random_transform(random_state)
    cone2origin[:3, 3] *= center_scale
    radius = min_radius + (1.0 - random_state.rand()) * radius_scale
    height = min_height + (1.0 - random_state.rand()) * height_scale
    return cone2origin, radius, height


def randn_convex(random_state, n_vertices=10, center_scale=1.0, min_radius=1.0,
                 radius_scale=1.0, return_triangles=True):
    """Random points on an ellipsoid surface (all of them hull vertices)."""
    phis = random_state.rand(n_vertices) * np.pi
    thetas = random_state.rand(n_vertices) * 2 * np.pi
    sin_phis = np.sin(phis)
    radii = min_radius + (1.0 - random_state.rand(3)) * radius_scale
    vertices = np.column_stack(
        (sin_phis * np.cos(thetas), sin_phis * np.sin(thetas), np.cos(phis))) * radii[np.newaxis]
    triangles = None
    if return_triangles:
        from scipy.spatial import ConvexHull
        triangles = ConvexHull(vertices - np.mean(vertices, axis=0)).simplices
    mesh2origin = random_transform(random_state)
    mesh2origin[:3, 3] *= center_scale
    return mesh2origin, vertices, triangles


def rand_ellipse(random_state, center_scale=1.0, radii_scale=1.0):
    return randn_rectangle(random_state, center_scale=center_scale, length_scale=radii_scale)


RANDOM_GENERATORS = {
    "sphere": rand_sphere,
    "ellipsoid": rand_ellipsoid,
    "capsule": rand_capsule,
    "disk": rand_circle,
    "ellipse": rand_ellipse,
    "cone": rand_cone,
    "cylinder": rand_cylinder,
    "box": rand_box,
    "mesh": randn_convex,
}


# ---------------------------------------------------------------------------
def random_transforms(rs, n):
    """n poses, exp of N(0, I6) exponential coordinates (vectorised)."""
    S = rs.randn(n, 6)
    w = S[:, :3]
    theta = np.linalg.norm(w, axis=1)
    safe = np.where(theta > 0.0, theta, 1.0)
    a = w / safe[:, None]
    v = S[:, 3:] / safe[:, None]
    K = np.zeros((n, 3, 3))
    K[:, 0, 1] = -a[:, 2]; K[:, 0, 2] = a[:, 1]
    K[:, 1, 0] = a[:, 2]; K[:, 1, 2] = -a[:, 0]
    K[:, 2, 0] = -a[:, 1]; K[:, 2, 1] = a[:, 0]
    K2 = K @ K
    s = np.sin(theta)[:, None, None]
    c = np.cos(theta)[:, None, None]
    th = theta[:, None, None]
    eye = np.eye(3)[None]
    R = eye + s * K + (1.0 - c) * K2
    V = eye * th + (1.0 - c) * K + (th - s) * K2
    T = np.zeros((n, 4, 4))
    T[:, :3, :3] = R
    T[:, :3, 3] = np.einsum("nij,nj->ni", V, v)
    T[:, 3, 3] = 1.0
    zero = theta == 0.0
    if np.any(zero):
        T[zero, :3, :3] = np.eye(3)
        T[zero, :3, 3] = S[zero, 3:]
    return T


PRIMITIVES = ("sphere", "ellipsoid", "capsule", "cylinder", "box")


def random_collider_set(rs, n, names=PRIMITIVES, center_scale=1.0, size_scale=1.0,
                        hull_vertices=(10, 10), hull_library=None, hull_min_radius=1.0):
    """n random colliders drawn uniformly from `names`, packed.

    Shape parameters follow the defaults of the reference generators
    (random.py:200-452): sizes / radii / heights uniform in (0, size_scale],
    centres and rotations from the pose distribution above scaled by
    `center_scale`.  "mesh" yields a world-frame `ConvexHullVertices` with a
    vertex count uniform in `hull_vertices` (inclusive), points on an ellipsoid
    with radii in (hull_min_radius, hull_min_radius + size_scale].
    `hull_library`: number of distinct hull shapes to draw from (None = every hull
    unique); poses are always unique.
    """
    code = {"sphere": _pack.SPHERE, "ellipsoid": _pack.ELLIPSOID, "capsule": _pack.CAPSULE,
            "cylinder": _pack.CYLINDER, "box": _pack.BOX, "mesh": _pack.HULL, "cone": _pack.CONE}
    codes = np.array([code[nm] for nm in names], dtype=np.int32)
    type_ = codes[rs.randint(len(names), size=n)]
    pose = random_transforms(rs, n)
    pose[:, :3, 3] *= center_scale
    param = (1.0 - rs.rand(n, 3)) * size_scale
    is_sphere = type_ == _pack.SPHERE
    pose[is_sphere, :3, :3] = np.eye(3)
    two = np.isin(type_, (_pack.SPHERE, _pack.CAPSULE, _pack.CYLINDER, _pack.CONE))
    param[two, 2] = 0.0
    param[is_sphere, 1] = 0.0
    vert_len = np.zeros(n, dtype=np.int64)
    vert_len[type_ == _pack.BOX] = 8
    hull_idx = np.where(type_ == _pack.HULL)[0]
    nh = len(hull_idx)
    if nh:
        vert_len[hull_idx] = rs.randint(hull_vertices[0], hull_vertices[1] + 1, size=nh)
    vert_off = np.zeros(n, dtype=np.int64)
    vert_off[1:] = np.cumsum(vert_len[:-1])
    total = int(vert_len.sum())
    if total >= 2 ** 31:
        raise ValueError("vertex pool exceeds int32 offsets; shard the set")
    verts = np.zeros((total, 3))
    if nh:
        lens = vert_len[hull_idx]
        tot_h = int(lens.sum())
        owner = np.repeat(np.arange(nh), lens)
        phis = rs.rand(tot_h) * np.pi
        thetas = rs.rand(tot_h) * 2 * np.pi
        radii = hull_min_radius + (1.0 - rs.rand(nh, 3)) * size_scale
        if hull_library is not None and hull_library < nh:
            # re-use a library of shapes: hull i takes the radii and angle stream of shape i % L
            lib_of = np.arange(nh) % hull_library
            radii = radii[lib_of]
            seeds = rs.randint(0, 2 ** 31 - 1, size=hull_library)
            start = np.zeros(nh, dtype=np.int64)
            start[1:] = np.cumsum(lens[:-1])
            within = np.arange(tot_h) - start[owner]
            h = (seeds[lib_of][owner].astype(np.uint64) * np.uint64(2654435761)
                 + within.astype(np.uint64) * np.uint64(40503)) % np.uint64(2 ** 32)
            phis = (h.astype(np.float64) / 2 ** 32) * np.pi
            h2 = (h * np.uint64(1103515245) + np.uint64(12345)) % np.uint64(2 ** 32)
            thetas = (h2.astype(np.float64) / 2 ** 32) * 2 * np.pi
        sp = np.sin(phis)
        local = np.column_stack((sp * np.cos(thetas), sp * np.sin(thetas), np.cos(phis))) * radii[owner]
        Rm = pose[hull_idx][owner]
        world = np.einsum("nij,nj->ni", Rm[:, :3, :3], local) + Rm[:, :3, 3]
        pos = np.repeat(vert_off[hull_idx], lens) + (np.arange(tot_h) - np.repeat(
            np.concatenate(([0], np.cumsum(lens[:-1]))), lens))
        verts[pos] = world
        pose[hull_idx] = np.eye(4)
        param[hull_idx] = 0.0
    return _pack.ColliderSet(type_, pose, param, vert_off, vert_len, verts)


def random_meshgraph_set(rs, n_meshes, n_mesh_colliders, n_other, hull_vertices=(10, 30),
                          names=PRIMITIVES, center_scale=1.0):
    """`n_mesh_colliders` MeshGraph colliders that share `n_meshes` random convex meshes
    (randn_convex shapes with a vertex count uniform in `hull_vertices`, each collider with
    its own pose) followed by `n_other` colliders drawn from `names`, packed."""
    from . import colliders as C
    meshes = []
    for _ in range(n_meshes):
        nv = int(rs.randint(hull_vertices[0], hull_vertices[1] + 1))
        _, V, tri = randn_convex(rs, n_vertices=nv)
        meshes.append(C.MeshGraph(np.eye(4), V, tri))
    poses = random_transforms(rs, n_mesh_colliders)
    poses[:, :3, 3] *= center_scale
    cols = []
    for k in range(n_mesh_colliders):
        src = meshes[k % n_meshes]
        m = C.MeshGraph(poses[k], src.vertices, src.triangles)
        m._graph = src.graph_record()  # one adjacency record per mesh
        cols.append(m)
    mesh_set = _pack.pack_colliders(cols)
    if n_other == 0:
        return mesh_set
    return _pack.concat_sets([mesh_set, random_collider_set(rs, n_other, names=names,
                                                            center_scale=center_scale)])


# ---------------------------------------------------------------------------
# Device-side generation (torch): the same distributions as above, drawn in HBM.  Used for
# the large benchmark sets (16 M shapes would take minutes on the host); the streams differ
# from the numpy ones, the distributions do not.  A given (seed, device type) yields the same
# set on every GPU, which is how all ranks of a job hold identical replicas.
def random_transforms_device(gen, n, device):
    import torch
    S = torch.randn((n, 6), generator=gen, device=device, dtype=torch.float64)
    w = S[:, :3]
    theta = torch.linalg.norm(w, dim=1)
    safe = torch.where(theta > 0.0, theta, torch.ones_like(theta))
    a = w / safe[:, None]
    v = S[:, 3:] / safe[:, None]
    K = torch.zeros((n, 3, 3), device=device, dtype=torch.float64)
    K[:, 0, 1] = -a[:, 2]; K[:, 0, 2] = a[:, 1]
    K[:, 1, 0] = a[:, 2]; K[:, 1, 2] = -a[:, 0]
    K[:, 2, 0] = -a[:, 1]; K[:, 2, 1] = a[:, 0]
    K2 = K @ K
    s = torch.sin(theta)[:, None, None]
    c = torch.cos(theta)[:, None, None]
    th = theta[:, None, None]
    eye = torch.eye(3, device=device, dtype=torch.float64)[None]
    R = eye + s * K + (1.0 - c) * K2
    V = eye * th + (1.0 - c) * K + (th - s) * K2
    T = torch.zeros((n, 4, 4), device=device, dtype=torch.float64)
    T[:, :3, :3] = R
    T[:, :3, 3] = torch.einsum("nij,nj->ni", V, v)
    T[:, 3, 3] = 1.0
    return T


def random_collider_set_device(seed, n, names=PRIMITIVES, center_scale=1.0, size_scale=1.0,
                               hull_vertices=(10, 10), hull_library=None, hull_min_radius=1.0,
                               device=None, chunk=1 << 21):
    """Device-resident counterpart of :func:`random_collider_set`: returns
    :class:`~distance3d_b200.pack.DeviceColliders` (nothing touches the host).  With
    `hull_library` the hulls are re-posed copies of that many shapes; the world-frame vertex
    pool is still unique per hull (the reference's ConvexHullVertices is a world-frame list)."""
    import torch
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    f64 = dict(device=device, dtype=torch.float64)
    code = {"sphere": _pack.SPHERE, "ellipsoid": _pack.ELLIPSOID, "capsule": _pack.CAPSULE,
            "cylinder": _pack.CYLINDER, "box": _pack.BOX, "mesh": _pack.HULL, "cone": _pack.CONE}
    codes = torch.tensor([code[nm] for nm in names], dtype=torch.int32, device=device)
    type_ = codes[torch.randint(len(names), (n,), generator=gen, device=device)]
    pose = torch.empty((n, 4, 4), **f64)
    for s0 in range(0, n, chunk):
        pose[s0:s0 + chunk] = random_transforms_device(gen, min(chunk, n - s0), device)
    pose[:, :3, 3] *= center_scale
    param = (1.0 - torch.rand((n, 3), generator=gen, **f64)) * size_scale
    is_sphere = type_ == _pack.SPHERE
    pose[is_sphere, :3, :3] = torch.eye(3, **f64)
    two = (type_ == _pack.SPHERE) | (type_ == _pack.CAPSULE) | (type_ == _pack.CYLINDER) | (type_ == _pack.CONE)
    param[two, 2] = 0.0
    param[is_sphere, 1] = 0.0
    vert_len = torch.zeros(n, dtype=torch.int64, device=device)
    vert_len[type_ == _pack.BOX] = 8
    hull_idx = torch.nonzero(type_ == _pack.HULL).flatten()
    nh = int(hull_idx.numel())
    L = nh if hull_library is None else min(int(hull_library), nh)
    if nh:
        lib_len = torch.randint(hull_vertices[0], hull_vertices[1] + 1, (max(L, 1),), generator=gen, device=device)
        lib_of = torch.arange(nh, device=device) % max(L, 1)
        vert_len[hull_idx] = lib_len[lib_of]
    vert_off = torch.cumsum(vert_len, 0) - vert_len
    total = int(vert_len.sum().item())
    if total >= 2 ** 31:
        raise ValueError("vertex pool exceeds int32 offsets; shard the set")
    verts = torch.zeros((max(total, 1), 3), **f64)
    if nh:
        # library shapes: points on an ellipsoid (randn_convex), local frame
        lib_off = torch.cumsum(lib_len, 0) - lib_len
        lib_total = int(lib_len.sum().item())
        phis = torch.rand(lib_total, generator=gen, **f64) * np.pi
        thetas = torch.rand(lib_total, generator=gen, **f64) * (2 * np.pi)
        radii = hull_min_radius + (1.0 - torch.rand((max(L, 1), 3), generator=gen, **f64)) * size_scale
        owner_lib = torch.repeat_interleave(torch.arange(max(L, 1), device=device), lib_len)
        sp = torch.sin(phis)
        lib_local = torch.stack((sp * torch.cos(thetas), sp * torch.sin(thetas), torch.cos(phis)), dim=1) * radii[owner_lib]
        per = max(1, chunk // max(1, hull_vertices[1]))
        for h0 in range(0, nh, per):     # world-frame copies, a slab of hulls at a time
            hs = hull_idx[h0:h0 + per]
            lens = vert_len[hs]
            owner = torch.repeat_interleave(torch.arange(hs.numel(), device=device), lens)
            first = torch.cumsum(lens, 0) - lens
            within = torch.arange(int(lens.sum().item()), device=device) - first[owner]
            src = lib_off[lib_of[h0:h0 + per]][owner] + within
            Rm = pose[hs][owner]
            world = torch.einsum("nij,nj->ni", Rm[:, :3, :3], lib_local[src]) + Rm[:, :3, 3]
            verts[vert_off[hs][owner] + within] = world
        pose[hull_idx] = torch.eye(4, **f64)
        param[hull_idx] = 0.0
    return _pack.DeviceColliders.from_tensors(type_, pose, param, vert_off.to(torch.int32),
                                              vert_len.to(torch.int32), verts)


def random_capsules_device(seed, n, center_scale=2.0, radius_scale=0.1, height_scale=0.5, device=None):
    """BASELINE configs[1] (vis_capsules_benchmark.py:22-30 scaled up) generated in HBM."""
    import torch
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    pose = random_transforms_device(gen, n, device)
    pose[:, :3, 3] *= center_scale
    param = torch.zeros((n, 3), device=device, dtype=torch.float64)
    param[:, 0] = (1.0 - torch.rand(n, generator=gen, device=device, dtype=torch.float64)) * radius_scale
    param[:, 1] = (1.0 - torch.rand(n, generator=gen, device=device, dtype=torch.float64)) * height_scale
    type_ = torch.full((n,), _pack.CAPSULE, dtype=torch.int32, device=device)
    return _pack.DeviceColliders.from_tensors(type_, pose, param)


def device_set_to_host(dc, idx=None):
    """Host :class:`ColliderSet` with the colliders `idx` of a device set (all by default): the
    CPU oracle consumes exactly the numbers the kernels saw."""
    import torch
    if idx is None:
        idx = torch.arange(dc.n, device=dc.device)
    idx = torch.as_tensor(idx, device=dc.device).long()
    lens = dc.vert_len[idx].long()
    offs = dc.vert_off[idx].long()
    owner = torch.repeat_interleave(torch.arange(idx.numel(), device=dc.device), lens)
    first = torch.cumsum(lens, 0) - lens
    within = torch.arange(int(lens.sum().item()), device=dc.device) - first[owner]
    verts = dc.verts[offs[owner] + within] if within.numel() else dc.verts[:0]
    cs = _pack.ColliderSet(dc.type[idx].cpu().numpy(), dc.pose[idx].cpu().numpy(), dc.param[idx].cpu().numpy(),
                           first.cpu().numpy(), lens.cpu().numpy(), verts.cpu().numpy(),
                           None if dc.margin is None else dc.margin[idx].cpu().numpy())
    cs.boxes_prepared = True   # box vertices were generated on the device (d3d_prepare)
    return cs


def random_pairs(rs, n_colliders, n_pairs):
    """Random (i, j) index pairs, i != j."""
    a = rs.randint(n_colliders, size=n_pairs)
    b = (a + 1 + rs.randint(n_colliders - 1, size=n_pairs)) % n_colliders
    return np.stack((a, b), axis=1).astype(np.int32)
