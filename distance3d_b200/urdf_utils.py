"""URDF helpers (reference: distance3d/urdf_utils.py:7-118)."""
import re
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
    """Collision frames of the own, parent and child link per collision object
    (urdf_utils.py:39-64)."""
    whitelist = {}
    info = LinkInfo(tm)
    for obj in tm.collision_objects:
        link = info.link(obj.frame)
        whitelist[obj.frame] = (
            info.collision_frames_attached_to_link(link)
            + info.collision_frames_attached_to_link(info.parent_link(link))
            + info.collision_frames_attached_to_link(info.child_link(link)))
    return whitelist


class LinkInfo:
    """Link relations of a UrdfTransformManager (urdf_utils.py:67-118).

    Like the reference this relies on the insertion order of `tm.transforms`:
    frames attached to links are registered before any joint, so the last
    (child, parent) entry seen for a parent is its child LINK.
    """

    def __init__(self, tm):
        self.tm = tm
        self.parent_links = {}
        self.child_links = {}
        for child, parent in tm.transforms:
            self.parent_links[child] = parent
            self.child_links[parent] = child
        self.prog_match_link = re.compile(r"collision:(.*)\/.*")

    def link(self, frame):
        result = self.prog_match_link.match(frame)
        if result is None:
            warnings.warn(f"Couldn't extract link of collision object at frame '{frame}'")
            return None
        return result.group(1)

    def child_link(self, link_frame):
        return self.child_links.get(link_frame, None)

    def parent_link(self, link_frame):
        return self.parent_links.get(link_frame, None)

    def collision_frames_attached_to_link(self, link_frame):
        prog = re.compile(f"collision:{link_frame}" + r"\/.*")
        return [node for node in self.tm.nodes if isinstance(node, str) and prog.match(node)]
