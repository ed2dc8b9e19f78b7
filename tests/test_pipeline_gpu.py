"""GPU: full pipeline (LBVH broad phase -> GJK -> EPA) against the oracle."""
import numpy as np
import pytest

from distance3d_b200 import _lib, pipeline, random as d3random
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def test_pipeline_matches_oracle_stage_by_stage():
    rs = np.random.RandomState(41)
    cs = d3random.random_collider_set(rs, 6000, names=d3random.PRIMITIVES + ("mesh",),
                                      center_scale=6.0, hull_vertices=(8, 40))
    res = pipeline.collide(cs, shard=False)
    cand = res.candidates.cpu().numpy()
    A = O.aabb(cs)
    tree = O.Tree()
    tree.insert_aabbs(A)
    ref_pairs = tree.query(A)
    ref_cand = ref_pairs[ref_pairs[:, 0] < ref_pairs[:, 1]]
    assert res.n_overlaps == len(ref_cand) == len(cand)                 # every unordered pair once
    assert 2 * len(ref_cand) + len(cs) == len(ref_pairs)               # the reference lists both orders + (i, i)
    assert set(map(tuple, cand.tolist())) == set(map(tuple, ref_cand.tolist()))
    g = res.gjk.cpu()
    ref = O.gjk_distance(cs, cand, n_threads=O.max_threads())
    assert np.array_equal(g["dist"], ref["dist"])
    assert np.array_equal(g["closest_a"], ref["a"]) and np.array_equal(g["closest_b"], ref["b"])
    hits = res.hits.cpu().numpy()
    assert np.array_equal(hits, np.where(ref["dist"] == 0.0)[0]) and len(hits) > 50
    idx = res.epa_index.cpu().numpy()
    assert np.array_equal(idx, hits)
    e = res.epa.cpu()
    full = ref["n_points"][hits] == 4           # EPA is defined for 4-point simplices only
    assert np.all(e["status"][~full] == 8) and (~full).sum() > 0
    ref_e = O.epa(cs, cand[hits][full], ref["Y"][hits][full], n_threads=O.max_threads())
    assert np.array_equal(e["status"][full], ref_e["status"])
    ok = ref_e["status"] != 7
    assert np.array_equal(e["mtv"][full][ok], ref_e["mtv"][ok])
