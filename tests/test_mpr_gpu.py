"""GPU: MPR kernel against the oracle and the real reference's outputs (bit-exact)."""
import numpy as np
import pytest

from distance3d_b200 import mpr, gjk, random as d3random
from oracle import cpu_oracle as O
from util import load_golden

pytestmark = pytest.mark.gpu


def _cpu(out):
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}


def test_golden_mpr_vs_reference_outputs():
    cs, g = load_golden("mpr.npz")
    res = _cpu(mpr.mpr_batch(cs, g["pairs"]))
    assert np.array_equal(res["hit"], g["hit"])
    hit = g["hit"].astype(bool)
    assert np.max(np.abs(res["depth"][hit] - g["depth"][hit])) < 1e-7
    assert np.array_equal(res["depth"][hit], g["depth"][hit])
    assert np.array_equal(res["dir"][hit], g["dir"][hit])
    assert np.array_equal(res["pos"][hit], g["pos"][hit])
    res_i = _cpu(mpr.mpr_batch(cs, g["pairs"], penetration=False))
    assert np.array_equal(res_i["hit"], g["hit_intersection"])


@pytest.mark.parametrize("names,scale", [(d3random.PRIMITIVES, 0.5), (d3random.PRIMITIVES + ("mesh", "cone"), 1.0)])
def test_random_vs_oracle_and_gjk(names, scale):
    rs = np.random.RandomState(31)
    cs = d3random.random_collider_set(rs, 2000, names=names, center_scale=scale)
    pairs = d3random.random_pairs(rs, len(cs), 30000)
    res = _cpu(mpr.mpr_batch(cs, pairs))
    ref = O.mpr(cs, pairs, n_threads=O.max_threads())
    assert np.array_equal(res["hit"], ref["hit"])
    assert np.array_equal(res["status"], ref["status"])
    hit = ref["hit"].astype(bool)
    assert np.array_equal(res["depth"][hit], ref["depth"][hit])
    assert np.array_equal(res["dir"][hit], ref["dir"][hit])
    assert np.array_equal(res["pos"][hit], ref["pos"][hit])
    # reference test_mpr.py:7-20: MPR and GJK agree on intersection (away from contact)
    g = gjk.gjk_distance_batch(cs, pairs).cpu()
    clear = (g["dist"] > 1e-3) | (g["dist"] == 0.0)
    agree = (res["hit"][clear] == 1) == (g["dist"][clear] == 0.0)
    assert agree.mean() > 0.995


def test_scalar_api_known_answers():
    # distance3d/test/test_mpr.py:23-82
    from distance3d_b200 import colliders as C
    s1 = C.Sphere(np.zeros(3), 1.0)
    hit, depth, d, pos = mpr.mpr_penetration(s1, C.Sphere(np.array([0.0, 0.0, 1.0]), 0.5))
    assert hit and depth == 0.5
    np.testing.assert_allclose(d, [0, 0, 1], atol=1e-6)
    np.testing.assert_allclose(pos, [0, 0, 0.75], atol=1e-6)
    assert mpr.mpr_penetration(s1, C.Sphere(np.array([0.0, 0.0, 5.0]), 0.5)) == (False, None, None, None)
    assert mpr.mpr_intersection(s1, s1) is True
