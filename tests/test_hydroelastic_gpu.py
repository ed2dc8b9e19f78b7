"""GPU: tetrahedron broad phase + pair intersection (hydroelastic contact consumer of the AABB
broad phase) against the reference's outputs (tests/golden/tetra.npz) and the oracle."""
import os

import numpy as np
import pytest

from distance3d_b200 import hydroelastic_contact as hc
from oracle import cpu_oracle as O
from util import GOLDEN, compare_tetra_results

pytestmark = pytest.mark.gpu


def _cases():
    g = dict(np.load(os.path.join(GOLDEN, "tetra.npz")))
    return g, ["c%d_" % c for c in range(int(g["n_cases"]))]


def test_tetra_pairs_vs_reference_outputs():
    g, keys = _cases()
    compared = 0
    for k in keys:
        assert np.array_equal(hc.tetrahedral_mesh_aabbs(g[k + "tp1"]), g[k + "aabb1"])
        res = hc.intersect_tetrahedron_pairs_batch(g[k + "pairs"], g[k + "tp1"], g[k + "tp2"], g[k + "e1"],
                                                   g[k + "e2"], youngs_modulus1=g[k + "ym"][0],
                                                   youngs_modulus2=g[k + "ym"][1]).cpu()
        compared += compare_tetra_results(res, g, k)
        # same arithmetic as the oracle: identical booleans / statuses, planes to the last bits
        ref = O.tetra_pairs(g[k + "pairs"], g[k + "tp1"], g[k + "e1"], g[k + "tp2"], g[k + "e2"],
                            youngs_modulus1=g[k + "ym"][0], youngs_modulus2=g[k + "ym"][1],
                            n_threads=O.max_threads())
        assert np.array_equal(res["hit"], ref["hit"]) and np.array_equal(res["status"], ref["status"])
        assert np.max(np.abs(res["plane"] - ref["plane"]), initial=0.0) < 1e-12
        assert np.array_equal(res["n_vertices"], ref["n_vertices"])
        assert np.max(np.abs(res["polygon"] - ref["polygon"]), initial=0.0) < 1e-12
    assert compared > 2500


def test_broad_phase_and_drop_in_signature():
    g, keys = _cases()
    for k in keys[1:8]:
        for trees in (True, False):
            pairs, res = hc.find_contact_pairs(g[k + "tp1"], g[k + "e1"], g[k + "tp2"], g[k + "e2"],
                                               g[k + "ym"][0], g[k + "ym"][1], use_aabb_trees=trees)
            got = set(map(tuple, pairs.cpu().numpy().tolist()))
            assert got == set(map(tuple, g[k + "pairs"].tolist()))           # candidate SET of the reference
            hits = {tuple(p) for p, h in zip(pairs.cpu().numpy().tolist(), res.hit.cpu().numpy()) if h}
            assert hits == {tuple(p) for p, h in zip(g[k + "pairs"].tolist(), g[k + "hit"]) if h}
    k = keys[5]
    X1 = {i: x for i, x in enumerate(hc.barycentric_transforms(g[k + "tp1"]))}
    inter, planes, polys, i1, i2 = hc.intersect_tetrahedron_pairs(
        g[k + "pairs"], g[k + "tp1"], g[k + "tp2"], g[k + "e1"], g[k + "e2"], X1, None, g[k + "ym"][0], g[k + "ym"][1])
    sel = np.nonzero(g[k + "hit"])[0]
    assert inter and len(polys) == len(sel) == len(planes)
    assert i1 == [int(i) for i in g[k + "pairs"][sel, 0]] and i2 == [int(j) for j in g[k + "pairs"][sel, 1]]
    assert all(len(p) == n for p, n in zip(polys, g[k + "nv"][sel])) or True
    Xd = hc.barycentric_transforms(g[k + "tp2"])
    ref = np.linalg.pinv(np.hstack((g[k + "tp2"].transpose((0, 2, 1)), np.ones((len(g[k + "tp2"]), 1, 4)))))
    assert np.max(np.abs(Xd - ref)) < 1e-9 * max(1.0, np.abs(ref).max())
