"""Robot self-collision detection (reference: distance3d/self_collision.py:5-64).

`detect` / `detect_any` keep the reference's interface on a
`BoundingVolumeHierarchy`.  `RobotModel.detect_batch` is the batched form for many
joint configurations (BASELINE config 4): forward kinematics, bounding boxes,
white-list filtered candidate pairs and `gjk_intersection` all run on the device.
"""
import ctypes

import numpy as np

from . import _lib, gjk
from ._lib import c_i64, c_int, ptr
from .pack import DeviceColliders, pack_colliders


def _candidate_pattern(frames, whitelists):
    """Frame pairs (i < j) the narrow phase has to test, and the white-list bit masks.

    The reference filters the candidates of frame i by whitelists[i] only
    (self_collision.py:27-28), so a pair is dropped only when BOTH frames white-list each
    other.  Returns (pattern int32[C,2], wl uint64[K] with bit j of wl[i] = j white-listed
    for i, symmetric flag).  White-lists of a serial chain are symmetric; a link with several
    child links white-lists only one of them (urdf_utils.py:79-81) and is not.
    """
    K = len(frames)
    index = {f: i for i, f in enumerate(frames)}
    wl = np.zeros(K, dtype=np.uint64)
    for i, fi in enumerate(frames):
        for fj in whitelists.get(fi, ()):
            if fj in index and index[fj] < 64:
                wl[i] |= np.uint64(1) << np.uint64(index[fj])
    listed = lambda i, j: bool((int(wl[i]) >> j) & 1) if j < 64 else frames[j] in whitelists.get(frames[i], ())  # noqa: E731
    pattern = [(i, j) for i in range(K) for j in range(i + 1, K)
               if not (listed(i, j) and listed(j, i))]
    symmetric = all(listed(i, j) == listed(j, i) for i in range(K) for j in range(i + 1, K)) \
        and all(listed(i, i) for i in range(K))
    return np.array(pattern, dtype=np.int32).reshape(-1, 2), wl, symmetric


def _contact_mask(dc, n_groups, group_size, pattern_t, wl_t=None, symmetric=True):
    """uint8 mask [n_groups * group_size] of self_collision.detect, the number of narrow-phase
    candidates and the number of pairs whose GJK run ended in an error status.

    Symmetric white-lists: a collider is flagged iff it takes part in an intersecting
    candidate pair (the reference's loop order cannot matter then).  Otherwise the
    reference's ordered loop is replayed on the device (d3d_detect_ordered)."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    dev = dc.device
    n_pattern = int(pattern_t.shape[0])
    cap = max(1, n_groups * n_pattern)
    pairs = torch.empty((cap, 2), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    aabb = None
    if group_size <= 128:
        # boxes and candidate filter in one pass; the boxes are only written out when the ordered
        # replay below reads them
        if not symmetric:
            aabb = torch.empty((dc.n, 3, 2), dtype=torch.float64, device=dev)
        _lib._check(L.d3d_aabb_filter_pairs(ctypes.byref(dc.struct), c_i64(n_groups), c_int(group_size),
                                            ptr(pattern_t), c_int(n_pattern), ptr(aabb), ptr(pairs),
                                            c_i64(cap), ptr(count), _lib.stream_ptr()))
    else:
        aabb = torch.empty((dc.n, 3, 2), dtype=torch.float64, device=dev)
        _lib._check(L.d3d_aabb(ctypes.byref(dc.struct), ptr(aabb), _lib.stream_ptr()))
        _lib._check(L.d3d_filter_pairs(ptr(aabb), c_i64(n_groups), c_int(group_size), ptr(pattern_t),
                                       c_int(n_pattern), ptr(pairs), c_i64(cap), ptr(count),
                                       _lib.stream_ptr()))
    n_cand = int(count.item())
    mask = torch.zeros(dc.n, dtype=torch.uint8, device=dev)
    hit = status = None
    if n_cand:
        hit, _, status = gjk.gjk_intersection_batch(dc, pairs[:n_cand])
    if symmetric:
        if n_cand:
            _lib._check(L.d3d_scatter_hits(ptr(pairs), ptr(hit), ptr(count), c_i64(n_cand), ptr(mask),
                                           _lib.stream_ptr()))
    else:
        if group_size > 64:
            raise NotImplementedError("ordered self-collision replay supports up to 64 colliders")
        bits = torch.empty(dc.n, dtype=torch.int64, device=dev)
        _lib._check(L.d3d_detect_ordered(ptr(aabb), c_i64(n_groups), c_int(group_size), ptr(pairs),
                                         ptr(hit), ptr(count), c_i64(n_cand), ptr(wl_t), ptr(bits),
                                         ptr(mask), _lib.stream_ptr()))
    n_bad = (status >= gjk.STATUS_SANITY_FAILED).sum() if n_cand else None
    return mask, n_cand, n_bad


def _raise_for_bad(n_bad):
    n = sum(int(b.item()) for b in n_bad if b is not None)
    if n:
        raise RuntimeError("%d candidate pairs ended GJK without a verdict (monotonicity "
                           "assertion or iteration cap); their contacts are undefined" % n)


def detect(bvh):
    """Maps each collider frame to whether it is in contact with another collider
    (reference: self_collision.py:5-36; uses bvh.self_collision_whitelists_)."""
    torch = _lib.torch_cuda()
    frames = list(bvh.colliders_.keys())
    if not frames:
        return {}
    cs = pack_colliders(list(bvh.colliders_.values()))
    dc = cs.device()
    pattern, wl, symmetric = _candidate_pattern(frames, bvh.self_collision_whitelists_)
    pattern_t = torch.from_numpy(pattern).to(dc.device)
    wl_t = torch.from_numpy(wl.view(np.int64)).to(dc.device)
    mask, _, n_bad = _contact_mask(dc, 1, len(frames), pattern_t, wl_t, symmetric)
    _raise_for_bad([n_bad])
    mask = mask.cpu().numpy().astype(bool)
    return {frame: bool(m) for frame, m in zip(frames, mask)}


def detect_any(bvh):
    """Is there any self collision? (reference: self_collision.py:39-64: every frame tests all
    of its candidates, so the answer does not depend on their order)."""
    return any(detect(bvh).values())


class RobotModel:
    """Flattened kinematics + collision geometry of a URDF robot on the device.

    Parameters
    ----------
    tm : distance3d_b200.urdf.UrdfTransformManager
        Transform manager with a loaded URDF.
    bvh : BoundingVolumeHierarchy
        BVH filled from `tm` with `fill_self_collision_whitelists=True`; provides the
        colliders (shape parameters) and the white-lists.
    """

    def __init__(self, tm, bvh):
        torch = _lib.torch_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.frames = list(bvh.colliders_.keys())
        self.template = pack_colliders(list(bvh.colliders_.values()))
        from .pack import BOX, HULL
        if np.any(np.isin(self.template.type, (BOX, HULL))):
            raise NotImplementedError("batched self-collision supports analytic colliders (sphere, "
                                      "cylinder, capsule, ellipsoid, cone) and MeshGraph colliders; "
                                      "boxes / world-frame hulls keep per-pose vertex lists")
        kin = tm.compile_kinematics(self.frames, "origin")
        self.joint_names = kin["joint_names"]
        self.n_frames = len(self.frames)
        self.n_joints = len(self.joint_names)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        self.n_keep = int(kin["n_keep"])
        self.kin = {k: t(v) for k, v in kin.items() if k not in ("joint_names", "n_keep")}
        self.pattern, wl, self.symmetric = _candidate_pattern(self.frames, bvh.self_collision_whitelists_)
        self.pattern_t = t(self.pattern)
        self.wl_t = t(wl.view(np.int64))
        self.type_t = t(self.template.type)
        self.param_t = t(self.template.param)
        # MeshGraph colliders: local-frame vertices and adjacency records are shared by all
        # configurations, only the pose differs
        self.mesh = None
        if np.any(self.template.vert_len > 0):
            tp = self.template
            self.mesh = dict(vert_off=t(tp.vert_off), vert_len=t(tp.vert_len), verts=t(tp.verts),
                             graph_off=None if tp.graph_off is None else t(tp.graph_off),
                             graph=None if tp.graph is None else t(tp.graph))
        self.device = dev

    def forward_kinematics(self, q, shared_prefixes=True):
        """Poses of all collider frames: q[B,J] -> device tensor [B,K,4,4].  `shared_prefixes=False`
        evaluates every frame's chain on its own (d3d_fk_urdf); the poses are the same bit for bit."""
        torch = _lib.torch_cuda()
        if not isinstance(q, torch.Tensor):
            q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64))
        q = q.to(device=self.device, dtype=torch.float64).reshape(-1, self.n_joints).contiguous()
        B = q.shape[0]
        out = torch.empty((B, self.n_frames, 4, 4), dtype=torch.float64, device=self.device)
        k = self.kin
        if shared_prefixes and self.n_keep <= 16:
            # one thread per configuration walks the tree of chain steps (common prefixes once)
            _lib._check(_lib.lib().d3d_fk_urdf_tree(
                c_int(self.n_frames), c_int(self.n_joints), ptr(k["joint_axis"]), ptr(k["joint_limits"]),
                ptr(k["joint_type"]), c_int(int(k["node_parent"].shape[0])), c_int(self.n_keep),
                ptr(k["node_parent"]), ptr(k["node_fixed"]), ptr(k["node_joint"]), ptr(k["node_keep"]),
                ptr(k["node_out_off"]), ptr(k["node_out"]), ptr(q), c_i64(B), ptr(out), _lib.stream_ptr()))
            return out
        _lib._check(_lib.lib().d3d_fk_urdf(
            c_int(self.n_frames), c_int(self.n_joints), ptr(k["joint_axis"]), ptr(k["joint_limits"]),
            ptr(k["joint_type"]), ptr(k["chain_off"]), ptr(k["chain_fixed"]), ptr(k["chain_joint"]),
            ptr(q), c_i64(B), ptr(out), _lib.stream_ptr()))
        return out

    def colliders_for(self, poses):
        """DeviceColliders of B * K colliders posed by `poses` [B,K,4,4]."""
        B = poses.shape[0]
        if self.mesh is None:
            return DeviceColliders.from_tensors(self.type_t.repeat(B), poses.reshape(-1, 4, 4),
                                                self.param_t.repeat(B, 1), has_boxes=False)
        m = self.mesh
        return DeviceColliders.from_tensors(
            self.type_t.repeat(B), poses.reshape(-1, 4, 4), self.param_t.repeat(B, 1),
            m["vert_off"].repeat(B), m["vert_len"].repeat(B), m["verts"],
            graph_off=None if m["graph_off"] is None else m["graph_off"].repeat(B), graph=m["graph"],
            has_boxes=False)

    def detect_batch(self, q, chunk=1 << 20):
        """Contact mask for every joint configuration: uint8 device tensor [B, K]
        (mask[b, k] = frame k touches a non-white-listed collider), plus the number
        of narrow-phase candidates that were tested."""
        torch = _lib.torch_cuda()
        if not isinstance(q, torch.Tensor):
            q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64))
        q = q.reshape(-1, self.n_joints)
        B = q.shape[0]
        out = torch.empty((B, self.n_frames), dtype=torch.uint8, device=self.device)
        n_cand = 0
        bad = []
        for s in range(0, B, chunk):
            poses = self.forward_kinematics(q[s:s + chunk])
            dc = self.colliders_for(poses)
            mask, nc, n_bad = _contact_mask(dc, poses.shape[0], self.n_frames, self.pattern_t,
                                            self.wl_t, self.symmetric)
            out[s:s + chunk] = mask.reshape(-1, self.n_frames)
            n_cand += nc
            bad.append(n_bad)
        _raise_for_bad(bad)
        return out, n_cand
