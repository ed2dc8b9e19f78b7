"""Broad-phase collision detection (reference: distance3d/broad_phase.py:13-262).

`BoundingVolumeHierarchy` keeps the reference's interface: a dict frame ->
collider plus an AABB tree whose external data are ``(frame, collider)`` tuples.
All bounding boxes of the scene are produced by ONE `d3d_aabb` launch over the
packed collider set and the tree is a device LBVH.
"""
import warnings

import numpy as np

from . import _lib
from .aabb_tree import AabbTree
from .colliders import Sphere, Box, Cylinder, MeshGraph
from .io import load_mesh
from .pack import pack_colliders
from . import urdf
from .urdf_utils import self_collision_whitelists


class BoundingVolumeHierarchy:
    """BVH over colliders attached to frames of a transform manager.

    Parameters and attributes as in the reference (broad_phase.py:13-51):
    `aabbtree_`, `colliders_`, `self_collision_whitelists_`.
    """

    def __init__(self, tm, base_frame, base_frame2origin=np.eye(4)):
        tm.add_transform(base_frame, "origin", base_frame2origin)
        self.tm = tm
        self.base_frame = base_frame
        self.base_frame2origin = base_frame2origin
        self.collider_frames = set()
        self.colliders_ = {}
        self.self_collision_whitelists_ = {}
        self._tree = None

    # -- tree over the current colliders (rebuilt lazily) ------------------
    @property
    def aabbtree_(self):
        if self._tree is None:
            tree = AabbTree()
            if self.colliders_:
                cs = pack_colliders(list(self.colliders_.values()))
                tree.insert_aabbs(_lib.aabb(cs), list(self.colliders_.items()))
            self._tree = tree
        return self._tree

    def packed(self):
        """Packed ColliderSet of all colliders (dict order) and the frame list."""
        return pack_colliders(list(self.colliders_.values())), list(self.colliders_.keys())

    def fill_tree_with_colliders(self, tm, make_artists=False,
                                 fill_self_collision_whitelists=False, use_visuals=False):
        """Fill the tree from a URDF transform manager (broad_phase.py:53-91)."""
        objects = tm.visuals if use_visuals else tm.collision_objects
        for obj in objects:
            try:
                collider = self._make_collider(tm, obj, make_artists)
                self.add_collider(obj.frame, collider)
            except RuntimeError as e:
                warnings.warn(str(e))
        if fill_self_collision_whitelists:
            self.self_collision_whitelists_.update(self_collision_whitelists(tm))
        self.update_collider_poses()

    def _make_collider(self, tm, obj, make_artists):
        """Collider for a URDF geometry (broad_phase.py:93-127)."""
        A2B = tm.get_transform(obj.frame, "origin")
        if isinstance(obj, urdf.Sphere):
            return Sphere(center=A2B[:3, 3], radius=obj.radius)
        if isinstance(obj, urdf.Box):
            return Box(A2B, obj.size)
        if isinstance(obj, urdf.Cylinder):
            return Cylinder(cylinder2origin=A2B, radius=obj.radius, length=obj.length)
        assert isinstance(obj, urdf.Mesh)
        if obj.filename is None:
            raise RuntimeError("mesh collider in frame '%s' has no file (no mesh_path / package_dir "
                               "was given to load_urdf)" % obj.frame)
        try:
            vertices, triangles = load_mesh(obj.filename, obj.scale)
        except (OSError, ValueError, IndexError) as e:  # unreadable or malformed file: a warning
            # (fill_tree_with_colliders), the other colliders of the robot are still loaded
            raise RuntimeError("mesh collider '%s' could not be loaded: %s" % (obj.filename, e))
        if len(vertices) == 0 or len(triangles) == 0:
            raise RuntimeError("mesh collider '%s' has no triangles" % obj.filename)
        return MeshGraph(A2B, vertices, triangles)

    def add_collider(self, frame, collider):
        """Add a collider located in `frame` (broad_phase.py:129-142)."""
        self.collider_frames.add(frame)
        self.colliders_[frame] = collider
        self._tree = None

    def update_collider_poses(self):
        """Pull all collider poses from the transform manager (broad_phase.py:144-151)."""
        for frame, collider in self.colliders_.items():
            collider.update_pose(self.tm.get_transform(frame, "origin"))
        self._tree = None

    def get_colliders(self):
        return self.colliders_.values()

    def get_artists(self):
        return [c.artist_ for c in self.colliders_.values() if c.artist_ is not None]

    def get_collider_frames(self):
        return self.collider_frames

    def aabb_overlapping_colliders(self, collider, whitelist=()):
        """Colliders whose AABB overlaps the AABB of `collider` (broad_phase.py:174-200)."""
        tree = self.aabbtree_
        _, overlaps = tree.overlaps_aabb(collider.aabb())
        colliders = dict(tree.external_data_list[int(i)] for i in overlaps)
        for frame in whitelist:
            colliders.pop(frame, None)
        return colliders

    def aabb_overlapping_with_other_bvh(self, other_bvh):
        """Pairs ((frame, collider), (frame, collider)) with overlapping AABBs
        (broad_phase.py:202-227)."""
        t1, t2 = self.aabbtree_, other_bvh.aabbtree_
        _, _, _, pairs = t1.overlaps_aabb_tree(t2)
        return [(t1.external_data_list[int(i)], t2.external_data_list[int(j)]) for i, j in pairs]

    def aabb_overlapping_with_self(self):
        """As above for the BVH with itself, without (i, i) (broad_phase.py:229-252)."""
        tree = self.aabbtree_
        _, _, _, pairs = tree.overlaps_aabb_tree(tree)
        return [(tree.external_data_list[int(i)], tree.external_data_list[int(j)])
                for i, j in pairs if i != j]
