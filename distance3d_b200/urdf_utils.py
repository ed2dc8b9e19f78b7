"""URDF helpers (reference: distance3d/urdf_utils.py:7-118)."""
import warnings

import numpy as np


def fast_transform_manager_initialization(tm, frames, base):
    """Register many frames defined w.r.t. `base` at once (urdf_utils.py:7-36)."""
    if base not in tm.nodes:
        tm.nodes.append(base)
    base_index = tm.nodes.index(base)
    for frame in frames:
        tm.nodes.append(frame)
        key = (frame, base)
        tm.transform_to_ij_index[key] = len(tm.i)
        tm.i.append(len(tm.nodes) - 1)
        tm.j.append(base_index)
        tm.transforms[key] = np.eye(4)
    tm._recompute_shortest_path()


def self_collision_whitelists(tm):
    """Frames a collision object may touch: every collision object of its own link, of the
    parent link and of the child link (urdf_utils.py:39-64)."""
    info = LinkInfo(tm)
    whitelist = {}
    for obj in tm.collision_objects:
        link = info.link(obj.frame)
        whitelist[obj.frame] = [
            frame for l in (link, info.parent_link(link), info.child_link(link))
            for frame in info.collision_frames_attached_to_link(l)]
    return whitelist


class LinkInfo:
    """Link relations read off the transform graph of a UrdfTransformManager.

    Every edge `(child, parent)` of `tm.transforms` is either a collision / visual /
    inertial frame hanging off its link or a joint between two links.  The collision frames
    are known (`tm.collision_objects`), so their edge names the owning link directly; the
    remaining edges give each link its parent and its children.  One behaviour of the
    reference (urdf_utils.py:67-118) is kept on purpose because the white-lists and with
    them the self-collision masks depend on it: a link with several child links white-lists
    only the LAST one registered (the reference overwrites `child_links[parent]` edge by
    edge, joints come after the link-attached frames).
    """

    def __init__(self, tm):
        self.tm = tm
        collision_frames = {obj.frame for obj in tm.collision_objects}
        self.owner = {}        # collision frame -> link
        self.attached = {}     # link -> its collision frames, registration order
        self.parent_links = {}
        self.child_links = {}  # link -> last registered child
        for child, parent in tm.transforms:
            if child in collision_frames:
                self.owner[child] = parent
                self.attached.setdefault(parent, []).append(child)
            self.parent_links[child] = parent
            self.child_links[parent] = child

    def link(self, frame):
        if frame not in self.owner:
            warnings.warn(f"Couldn't extract link of collision object at frame '{frame}'")
            return None
        return self.owner[frame]

    def child_link(self, link_frame):
        return self.child_links.get(link_frame)

    def parent_link(self, link_frame):
        return self.parent_links.get(link_frame)

    def collision_frames_attached_to_link(self, link_frame):
        return list(self.attached.get(link_frame, ()))
