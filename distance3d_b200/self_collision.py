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
    """Unordered frame pairs (i < j) that are not white-listed in either direction."""
    pattern = []
    for i, fi in enumerate(frames):
        for j in range(i + 1, len(frames)):
            fj = frames[j]
            if fj in whitelists.get(fi, ()) or fi in whitelists.get(fj, ()):
                continue
            pattern.append((i, j))
    return np.array(pattern, dtype=np.int32).reshape(-1, 2)


def _contact_mask(dc, n_groups, group_size, pattern_t):
    """uint8 mask [n_groups * group_size]: collider takes part in an intersecting candidate pair."""
    torch = _lib.torch_cuda()
    L = _lib.lib()
    dev = dc.device
    aabb = torch.empty((dc.n, 3, 2), dtype=torch.float64, device=dev)
    _lib._check(L.d3d_aabb(ctypes.byref(dc.struct), ptr(aabb), _lib.stream_ptr()))
    n_pattern = int(pattern_t.shape[0])
    cap = max(1, n_groups * n_pattern)
    pairs = torch.empty((cap, 2), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib._check(L.d3d_filter_pairs(ptr(aabb), c_i64(n_groups), c_int(group_size), ptr(pattern_t),
                                   c_int(n_pattern), ptr(pairs), c_i64(cap), ptr(count),
                                   _lib.stream_ptr()))
    n_cand = int(count.item())
    mask = torch.zeros(dc.n, dtype=torch.uint8, device=dev)
    if n_cand:
        hit, _, _ = gjk.gjk_intersection_batch(dc, pairs[:n_cand])
        _lib._check(L.d3d_scatter_hits(ptr(pairs), ptr(hit), ptr(count), c_i64(n_cand), ptr(mask),
                                       _lib.stream_ptr()))
    return mask, n_cand


def detect(bvh):
    """Maps each collider frame to whether it is in contact with another collider
    (reference: self_collision.py:5-36; uses bvh.self_collision_whitelists_)."""
    torch = _lib.torch_cuda()
    frames = list(bvh.colliders_.keys())
    if not frames:
        return {}
    cs = pack_colliders(list(bvh.colliders_.values()))
    dc = cs.device()
    pattern = _candidate_pattern(frames, bvh.self_collision_whitelists_)
    pattern_t = torch.from_numpy(pattern).to(dc.device)
    mask, _ = _contact_mask(dc, 1, len(frames), pattern_t)
    mask = mask.cpu().numpy().astype(bool)
    return {frame: bool(m) for frame, m in zip(frames, mask)}


def detect_any(bvh):
    """Is there any self collision? (reference: self_collision.py:39-64)."""
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
        if np.any(self.template.vert_len > 0):
            raise NotImplementedError("batched self-collision supports analytic colliders "
                                      "(sphere, cylinder, capsule, ellipsoid, cone) only")
        kin = tm.compile_kinematics(self.frames, "origin")
        self.joint_names = kin["joint_names"]
        self.n_frames = len(self.frames)
        self.n_joints = len(self.joint_names)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        self.kin = {k: t(v) for k, v in kin.items() if k != "joint_names"}
        self.pattern = _candidate_pattern(self.frames, bvh.self_collision_whitelists_)
        self.pattern_t = t(self.pattern)
        self.type_t = t(self.template.type)
        self.param_t = t(self.template.param)
        self.device = dev

    def forward_kinematics(self, q):
        """Poses of all collider frames: q[B,J] -> device tensor [B,K,4,4]."""
        torch = _lib.torch_cuda()
        if not isinstance(q, torch.Tensor):
            q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64))
        q = q.to(device=self.device, dtype=torch.float64).reshape(-1, self.n_joints).contiguous()
        B = q.shape[0]
        out = torch.empty((B, self.n_frames, 4, 4), dtype=torch.float64, device=self.device)
        k = self.kin
        _lib._check(_lib.lib().d3d_fk_urdf(
            c_int(self.n_frames), c_int(self.n_joints), ptr(k["joint_axis"]), ptr(k["joint_limits"]),
            ptr(k["joint_type"]), ptr(k["chain_off"]), ptr(k["chain_fixed"]), ptr(k["chain_joint"]),
            ptr(q), c_i64(B), ptr(out), _lib.stream_ptr()))
        return out

    def colliders_for(self, poses):
        """DeviceColliders of B * K colliders posed by `poses` [B,K,4,4]."""
        B = poses.shape[0]
        return DeviceColliders.from_tensors(self.type_t.repeat(B), poses.reshape(-1, 4, 4),
                                            self.param_t.repeat(B, 1))

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
        for s in range(0, B, chunk):
            poses = self.forward_kinematics(q[s:s + chunk])
            dc = self.colliders_for(poses)
            mask, nc = _contact_mask(dc, poses.shape[0], self.n_frames, self.pattern_t)
            out[s:s + chunk] = mask.reshape(-1, self.n_frames)
            n_cand += nc
        return out, n_cand
