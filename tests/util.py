"""Shared helpers for the tests: golden fixtures and random packed colliders."""
import os

import numpy as np

from distance3d_b200.pack import ColliderSet

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, name)))
    cs = ColliderSet(d["cs_type"], d["cs_pose"], d["cs_param"], d["cs_vert_off"], d["cs_vert_len"],
                     d["cs_verts"], d.get("cs_margin"), graph_off=d.get("cs_graph_off"),
                     graph=d.get("cs_graph"), mesh_start=d.get("cs_mesh_start"))
    return cs, d
