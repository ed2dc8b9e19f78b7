"""Minimal transform graph and URDF kinematics (host side, numpy).

The reference delegates forward kinematics to the un-vendored `pytransform3d`
package (reference call sites: distance3d/broad_phase.py:44,111,148,
distance3d/urdf_utils.py:21-36,79-118, tests distance3d/test/test_broad_phase.py:12-25,
distance3d/test/test_self_collision.py:10-41).  This module restates exactly the
semantics those call sites rely on:

* frames are hashable keys; ``add_transform(a, b, a2b)`` stores ``(a, b)``;
  ``get_transform(a, b)`` concatenates along the shortest path, inverting
  edges that are traversed backwards;
* URDF ``rpy`` is extrinsic x-y-z, a revolute joint's child-to-parent
  transform is ``origin * Rot(axis, q)``, ``q`` is clipped to ``<limit>``;
* collision frames are named ``collision:<link>/<name-or-index>``, visual
  frames ``visual:<link>/<name-or-index>``; all frames attached to links are
  registered before any joint transform and the root link is attached to a
  frame named after the robot.

It additionally compiles the kinematic tree into flat arrays for the batched
FK kernel (`d3d_fk_urdf`, see include/d3d_b200.h).
"""
import os
import xml.etree.ElementTree as ET
from collections import deque

import numpy as np

from ._transforms import (
    transform_from, active_matrix_from_extrinsic_euler_xyz,
    matrix_from_axis_angle, invert_transform)


class TransformManager:
    """Graph of rigid transformations between named frames."""

    def __init__(self, strict_check=True, check=True):
        self.strict_check = strict_check
        self.check = check
        self.nodes = []
        self.transforms = {}
        self.i = []
        self.j = []
        self.transform_to_ij_index = {}
        self._adj = {}

    def _recompute_shortest_path(self):
        self._adj = {}
        for (a, b) in self.transforms:
            self._adj.setdefault(a, []).append(b)
            self._adj.setdefault(b, []).append(a)

    def has_frame(self, frame):
        return frame in self.nodes

    def add_transform(self, from_frame, to_frame, A2B):
        A2B = np.asarray(A2B, dtype=float)
        if from_frame not in self.nodes:
            self.nodes.append(from_frame)
        if to_frame not in self.nodes:
            self.nodes.append(to_frame)
        key = (from_frame, to_frame)
        if key not in self.transforms:
            self.transform_to_ij_index[key] = len(self.i)
            self.i.append(self.nodes.index(from_frame))
            self.j.append(self.nodes.index(to_frame))
            self._adj.setdefault(from_frame, []).append(to_frame)
            self._adj.setdefault(to_frame, []).append(from_frame)
        self.transforms[key] = A2B
        return self

    def _path(self, from_frame, to_frame):
        if len(self._adj) == 0 and len(self.transforms) > 0:
            self._recompute_shortest_path()
        prev = {from_frame: None}
        queue = deque([from_frame])
        while queue:
            n = queue.popleft()
            if n == to_frame:
                break
            for m in self._adj.get(n, ()):
                if m not in prev:
                    prev[m] = n
                    queue.append(m)
        if to_frame not in prev:
            raise KeyError("Cannot compute path from frame '%s' to frame '%s'."
                           % (from_frame, to_frame))
        path = [to_frame]
        while prev[path[-1]] is not None:
            path.append(prev[path[-1]])
        return path[::-1]

    def get_transform(self, from_frame, to_frame):
        if from_frame not in self.nodes:
            raise KeyError("Unknown frame '%s'" % (from_frame,))
        if to_frame not in self.nodes:
            raise KeyError("Unknown frame '%s'" % (to_frame,))
        if (from_frame, to_frame) in self.transforms:
            return self.transforms[(from_frame, to_frame)]
        if (to_frame, from_frame) in self.transforms:
            return invert_transform(self.transforms[(to_frame, from_frame)])
        path = self._path(from_frame, to_frame)
        A2B = np.eye(4)
        for a, b in zip(path[:-1], path[1:]):
            if (a, b) in self.transforms:
                step = self.transforms[(a, b)]
            else:
                step = invert_transform(self.transforms[(b, a)])
            A2B = np.dot(step, A2B)
        return A2B


class Geometry:
    def __init__(self, frame, mesh_path=None, package_dir=None, color=None):
        self.frame = frame
        self.mesh_path = mesh_path
        self.package_dir = package_dir
        self.color = color


class Box(Geometry):
    def __init__(self, frame, mesh_path=None, package_dir=None, color=None):
        super().__init__(frame, mesh_path, package_dir, color)
        self.size = np.zeros(3)

    def parse(self, el):
        if "size" in el.attrib:
            self.size[:] = np.fromstring(el.attrib["size"], sep=" ")


class Sphere(Geometry):
    def __init__(self, frame, mesh_path=None, package_dir=None, color=None):
        super().__init__(frame, mesh_path, package_dir, color)
        self.radius = 0.0

    def parse(self, el):
        if "radius" not in el.attrib:
            raise ValueError("Sphere has no radius.")
        self.radius = float(el.attrib["radius"])


class Cylinder(Geometry):
    def __init__(self, frame, mesh_path=None, package_dir=None, color=None):
        super().__init__(frame, mesh_path, package_dir, color)
        self.radius = 0.0
        self.length = 0.0

    def parse(self, el):
        if "radius" not in el.attrib:
            raise ValueError("Cylinder has no radius.")
        self.radius = float(el.attrib["radius"])
        if "length" not in el.attrib:
            raise ValueError("Cylinder has no length.")
        self.length = float(el.attrib["length"])


class Mesh(Geometry):
    def __init__(self, frame, mesh_path=None, package_dir=None, color=None):
        super().__init__(frame, mesh_path, package_dir, color)
        self.filename = None
        self.scale = np.ones(3)

    def parse(self, el):
        if self.mesh_path is None and self.package_dir is None:
            self.filename = None
        else:
            if "filename" not in el.attrib:
                raise ValueError("Mesh has no filename.")
            if self.mesh_path is not None:
                self.filename = os.path.join(self.mesh_path, el.attrib["filename"])
            else:
                self.filename = el.attrib["filename"].replace(
                    "package://", self.package_dir)
            if "scale" in el.attrib:
                self.scale = np.fromstring(el.attrib["scale"], sep=" ")


_GEOMETRY = {"box": Box, "sphere": Sphere, "cylinder": Cylinder, "mesh": Mesh}


def _parse_origin(entry):
    """URDF <origin xyz rpy> -> 4x4 (rpy = extrinsic x-y-z Euler angles)."""
    origin = entry.find("origin") if entry is not None else None
    xyz = np.zeros(3)
    rpy = np.zeros(3)
    if origin is not None:
        if "xyz" in origin.attrib:
            xyz = np.fromstring(origin.attrib["xyz"], sep=" ")
        if "rpy" in origin.attrib:
            rpy = np.fromstring(origin.attrib["rpy"], sep=" ")
    return transform_from(active_matrix_from_extrinsic_euler_xyz(rpy), xyz)


class UrdfTransformManager(TransformManager):
    """Transform manager that loads URDF robots and articulates joints."""

    def __init__(self, strict_check=True, check=True):
        super().__init__(strict_check, check)
        self._joints = {}
        self.collision_objects = []
        self.visuals = []
        self.robot_name = None
        self.link_names = []

    def add_joint(self, joint_name, from_frame, to_frame, child2parent, axis,
                  limits=(float("-inf"), float("inf")), joint_type="revolute"):
        self.add_transform(from_frame, to_frame, child2parent)
        self._joints[joint_name] = (
            from_frame, to_frame, child2parent, np.asarray(axis, dtype=float),
            limits, joint_type)

    def set_joint(self, joint_name, value):
        if joint_name not in self._joints:
            raise KeyError("Joint '%s' is not known" % joint_name)
        from_frame, to_frame, child2parent, axis, limits, joint_type = \
            self._joints[joint_name]
        value = np.clip(value, limits[0], limits[1])
        if joint_type == "revolute":
            joint2A = transform_from(matrix_from_axis_angle(axis, value), np.zeros(3))
        else:
            joint2A = transform_from(np.eye(3), value * axis)
        self.add_transform(from_frame, to_frame, np.dot(child2parent, joint2A))

    def get_joint_limits(self, joint_name):
        return self._joints[joint_name][4]

    def load_urdf(self, urdf_xml, mesh_path=None, package_dir=None):
        root = ET.fromstring(urdf_xml)
        if root.tag != "robot":
            raise ValueError("Robot tag is missing.")
        if "name" not in root.attrib:
            raise ValueError("Attribute 'name' is missing in robot tag.")
        self.robot_name = root.attrib["name"]

        link_transforms = []
        self.link_names = []
        for link in root.findall("link"):
            name = link.attrib["name"]
            self.link_names.append(name)
            for kind, target in (("visual", self.visuals),
                                 ("collision", self.collision_objects)):
                for idx, entry in enumerate(link.findall(kind)):
                    entry_name = entry.attrib.get("name", idx)
                    frame = "%s:%s/%s" % (kind, name, entry_name)
                    link_transforms.append((frame, name, _parse_origin(entry)))
                    geometry = entry.find("geometry")
                    if geometry is None:
                        continue
                    for shape in geometry:
                        if shape.tag in _GEOMETRY:
                            obj = _GEOMETRY[shape.tag](frame, mesh_path, package_dir)
                            obj.parse(shape)
                            target.append(obj)
            inertial = link.find("inertial")
            if inertial is not None:
                link_transforms.append(
                    ("inertial_frame:%s" % name, name, _parse_origin(inertial)))

        self.add_transform(self.link_names[0], self.robot_name, np.eye(4))
        for t in link_transforms:
            self.add_transform(*t)

        for joint in root.findall("joint"):
            jname = joint.attrib["name"]
            jtype = joint.attrib["type"]
            parent = joint.find("parent").attrib["link"]
            child = joint.find("child").attrib["link"]
            child2parent = _parse_origin(joint)
            if jtype in ("revolute", "continuous", "prismatic"):
                axis_el = joint.find("axis")
                axis = np.array([1.0, 0.0, 0.0])
                if axis_el is not None and "xyz" in axis_el.attrib:
                    axis = np.fromstring(axis_el.attrib["xyz"], sep=" ")
                axis = axis / np.linalg.norm(axis)
                lower, upper = float("-inf"), float("inf")
                limit = joint.find("limit")
                if limit is not None:
                    lower = float(limit.attrib.get("lower", lower))
                    upper = float(limit.attrib.get("upper", upper))
                self.add_joint(
                    jname, child, parent, child2parent, axis, (lower, upper),
                    "prismatic" if jtype == "prismatic" else "revolute")
            else:
                self.add_transform(child, parent, child2parent)

    # ------------------------------------------------------------------
    # Flat kinematic model for the batched FK kernel
    # ------------------------------------------------------------------
    def compile_kinematics(self, frames, base_frame):
        """Flatten the kinematic chains of `frames` up to `base_frame`.

        Returns a dict of arrays describing, per frame, the product
        ``base<-...<-frame`` as an alternating sequence of fixed transforms
        and joint rotations:

        * ``joint_names``: list of the J actuated joints (dict order)
        * ``joint_axis  f64[J,3]``, ``joint_limits f64[J,2]``,
          ``joint_type int32[J]`` (0 revolute, 1 prismatic)
        * ``chain_off int32[K+1]``: range of chain steps per frame
        * ``chain_fixed f64[S,4,4]``: fixed transform applied at that step
        * ``chain_joint int32[S]``: joint index applied AFTER the fixed
          transform of the step (-1: none).  Pose = prod_s (fixed_s * rot_s),
          steps ordered from the base towards the frame.
        """
        joint_names = list(self._joints.keys())
        joint_of_edge = {}
        for jidx, (name, (frm, to, c2p, axis, limits, jt)) in enumerate(self._joints.items()):
            joint_of_edge[(frm, to)] = jidx
        chain_off = [0]
        chain_fixed = []
        chain_joint = []
        for frame in frames:
            path = self._path(frame, base_frame)  # frame ... base
            steps_fixed = []
            steps_joint = []
            # walk from base to frame: pose = T(base<-n1) T(n1<-n2) ... T(nk<-frame)
            for b, a in zip(path[::-1][:-1], path[::-1][1:]):
                # need a2b (child a expressed in b)
                if (a, b) in joint_of_edge:
                    jidx = joint_of_edge[(a, b)]
                    steps_fixed.append(self._joints[joint_names[jidx]][2])
                    steps_joint.append(jidx)
                elif (a, b) in self.transforms:
                    steps_fixed.append(self.transforms[(a, b)])
                    steps_joint.append(-1)
                elif (b, a) in joint_of_edge:
                    raise NotImplementedError(
                        "kinematic chain traverses a joint backwards")
                else:
                    steps_fixed.append(invert_transform(self.transforms[(b, a)]))
                    steps_joint.append(-1)
            # merge consecutive fixed transforms to shorten the chain
            merged_fixed = []
            merged_joint = []
            acc = np.eye(4)
            for T, j in zip(steps_fixed, steps_joint):
                acc = np.dot(acc, T)
                if j >= 0:
                    merged_fixed.append(acc)
                    merged_joint.append(j)
                    acc = np.eye(4)
            merged_fixed.append(acc)
            merged_joint.append(-1)
            chain_fixed.extend(merged_fixed)
            chain_joint.extend(merged_joint)
            chain_off.append(len(chain_fixed))
        # The steps of all chains as a tree (d3d_fk_urdf_tree): two chains share a node when they
        # agree in every step up to it, so the common prefix of the frames' chains is evaluated
        # once per configuration.  Nodes are created parents first.
        node_of = {}
        node_parent, node_fixed, node_joint, frame_node = [], [], [], []
        for k in range(len(frames)):
            parent = -1
            for s_ in range(chain_off[k], chain_off[k + 1]):
                key = (parent, np.ascontiguousarray(chain_fixed[s_], dtype=float).tobytes(), int(chain_joint[s_]))
                if key not in node_of:
                    node_of[key] = len(node_parent)
                    node_parent.append(parent)
                    node_fixed.append(chain_fixed[s_])
                    node_joint.append(int(chain_joint[s_]))
                parent = node_of[key]
            frame_node.append(parent)
        n_nodes = len(node_parent)
        has_child = np.zeros(n_nodes, dtype=bool)
        for p_ in node_parent:
            if p_ >= 0:
                has_child[p_] = True
        node_keep = np.full(n_nodes, -1, dtype=np.int32)
        node_keep[has_child] = np.arange(int(has_child.sum()), dtype=np.int32)
        order = np.argsort(np.asarray(frame_node, dtype=np.int64), kind="stable")
        node_out_off = np.zeros(n_nodes + 1, dtype=np.int32)
        np.add.at(node_out_off, np.asarray(frame_node, dtype=np.int64) + 1, 1)
        node_out_off = np.cumsum(node_out_off).astype(np.int32)
        J = len(joint_names)
        return {
            "node_parent": np.asarray(node_parent, dtype=np.int32),
            "node_fixed": np.array(node_fixed, dtype=float).reshape(-1, 4, 4),
            "node_joint": np.asarray(node_joint, dtype=np.int32),
            "node_keep": node_keep,
            "node_out_off": node_out_off,
            "node_out": order.astype(np.int32),
            "n_keep": int(has_child.sum()),
            "joint_names": joint_names,
            "joint_axis": np.array([self._joints[n][3] for n in joint_names]).reshape(J, 3),
            "joint_limits": np.array([self._joints[n][4] for n in joint_names], dtype=float).reshape(J, 2),
            "joint_type": np.array([0 if self._joints[n][5] == "revolute" else 1
                                    for n in joint_names], dtype=np.int32),
            "chain_off": np.array(chain_off, dtype=np.int32),
            "chain_fixed": np.array(chain_fixed, dtype=float).reshape(-1, 4, 4),
            "chain_joint": np.array(chain_joint, dtype=np.int32),
        }
