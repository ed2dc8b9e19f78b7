"""GPU: EPA (thread-per-pair kernel + warp kernel) against the oracle and the real reference's outputs.
`want_faces=True` and batches below 20000 pairs run the warp kernel alone; D3D_EPA_KERNEL=thread
forces the thread kernel (which hands over what it cannot finish) so that it is tested on the small
fixtures as well.

Tolerance from BASELINE.json: penetration depth and normal within 1e-7; the
implementation is in fact bit-exact (asserted)."""
import numpy as np
import pytest

from distance3d_b200 import epa, gjk, random as d3random
from oracle import cpu_oracle as O
from util import load_golden

pytestmark = pytest.mark.gpu


def test_golden_epa_vs_reference_outputs():
    cs, g = load_golden("epa.npz")
    sel = np.where(g["status"] >= 0)[0]
    res = epa.epa_batch(cs, g["pairs"][sel], g["Y"][sel], want_faces=True).cpu()
    asserted = g["status"][sel] == 7
    assert np.array_equal(res["status"] == 7, asserted)
    m = ~asserted
    depth = np.linalg.norm(res["mtv"][m], axis=1)
    ref_depth = np.linalg.norm(g["mtv"][sel][m], axis=1)
    assert np.max(np.abs(depth - ref_depth)) < 1e-7
    assert np.array_equal(res["mtv"][m], g["mtv"][sel][m])
    assert np.array_equal(res["success"][m], g["success"][sel][m])
    assert np.array_equal(res["n_faces"][m], g["n_faces"][sel][m])
    for q in np.where(m)[0]:
        n = res["n_faces"][q]
        assert np.array_equal(res["faces"][q, :n], g["faces"][sel[q], :n])


def test_golden_epa_thread_kernel(monkeypatch):
    monkeypatch.setenv("D3D_EPA_KERNEL", "thread")
    cs, g = load_golden("epa.npz")
    sel = np.where(g["status"] >= 0)[0]
    res = epa.epa_batch(cs, g["pairs"][sel], g["Y"][sel]).cpu()
    asserted = g["status"][sel] == 7
    assert np.array_equal(res["status"] == 7, asserted)
    m = ~asserted
    assert np.array_equal(res["mtv"][m], g["mtv"][sel][m])
    assert np.array_equal(res["success"][m], g["success"][sel][m])
    assert np.array_equal(res["n_faces"][m], g["n_faces"][sel][m])
    assert int(res["deferred"][0]) < 0.2 * len(sel)


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_golden_degenerate_simplices_vs_reference_outputs(monkeypatch, kernel):
    """epa_degenerate.npz: outputs of the real reference for simplices with a duplicated point or two
    points 1e-9 apart - shared vertex ids and near pairs in the thread kernel, coordinate matching in
    the warp kernel."""
    monkeypatch.setenv("D3D_EPA_KERNEL", kernel)
    cs, g = load_golden("epa_degenerate.npz")
    res = epa.epa_batch(cs, g["pairs"], g["Y"]).cpu()
    asserted = g["status"] == 7
    assert np.array_equal(res["status"] == 7, asserted)
    m = ~asserted
    assert np.array_equal(res["mtv"][m], g["mtv"][m])
    assert np.array_equal(res["success"][m], g["success"][m])
    assert np.array_equal(res["n_faces"][m], g["n_faces"][m])
    if kernel == "thread":
        assert int(res["deferred"][0]) < 0.1 * len(m)


def test_golden_wide_hulls():
    cs, g = load_golden("hulls.npz")
    sel = np.where(g["epa_status"] >= 0)[0]
    res = epa.epa_batch(cs, g["pairs"][sel], g["Y"][sel]).cpu()
    asserted = g["epa_status"][sel] == 7
    assert np.array_equal(res["status"] == 7, asserted)
    assert np.array_equal(res["mtv"][~asserted], g["epa_mtv"][sel][~asserted])
    assert np.array_equal(res["success"][~asserted], g["epa_success"][sel][~asserted])


@pytest.mark.parametrize("names,scale,hv", [
    (d3random.PRIMITIVES, 0.3, (10, 10)),
    (("mesh",), 0.6, (64, 256)),
    (("mesh", "box", "capsule"), 0.4, (8, 40)),
])
def test_random_pipeline_gjk_then_epa(names, scale, hv, monkeypatch):
    rs = np.random.RandomState(21)
    cs = d3random.random_collider_set(rs, 1500, names=names, center_scale=scale, hull_vertices=hv)
    pairs = d3random.random_pairs(rs, len(cs), 12000)
    g = gjk.gjk_distance_batch(cs, pairs)
    gc = g.cpu()
    sel = np.where((gc["dist"] == 0.0) & (gc["n_points"] == 4))[0]
    assert len(sel) > 500
    res = epa.epa_batch(cs, pairs[sel], g.simplex[sel.tolist()], want_faces=True).cpu()
    ref = O.epa(cs, pairs[sel], gc["simplex"][sel], return_faces=True, n_threads=O.max_threads())
    assert np.array_equal(res["status"], ref["status"])
    ok = ref["status"] != 7
    assert np.array_equal(res["mtv"][ok], ref["mtv"][ok])
    assert np.array_equal(res["success"][ok], ref["success"][ok])
    assert np.array_equal(res["n_faces"][ok], ref["n_faces"][ok])
    assert np.array_equal(res["iters"][ok], ref["iters"][ok])
    for q in np.where(ok)[0][:300]:
        n = ref["n_faces"][q]
        assert np.array_equal(res["faces"][q, :n], ref["faces"][q, :n])
    # property (reference test_epa.py:37-60): translating B by mtv separates the shapes
    conv = ok & (ref["success"] == 1)
    assert conv.sum() > 100
    # the same batch through the thread-per-pair kernel
    monkeypatch.setenv("D3D_EPA_KERNEL", "thread")
    thr = epa.epa_batch(cs, pairs[sel], g.simplex[sel.tolist()]).cpu()
    assert np.array_equal(thr["status"], ref["status"])
    for key in ("mtv", "success", "n_faces", "iters"):
        assert np.array_equal(thr[key][ok], ref[key][ok]), key
    if hv[0] > 32:   # hulls with more than 32 vertices are the warp kernel's (cooperative vertex scan)
        assert int(thr["deferred"][0]) == len(sel)
    else:
        assert int(thr["deferred"][0]) < 0.5 * len(sel)


def test_scalar_epa_known_answer():
    # distance3d/test/test_epa.py:7-34
    from distance3d_b200 import colliders as C
    from test_reference_kats import EPA_VERTICES1, EPA_VERTICES2
    c1, c2 = C.ConvexHullVertices(EPA_VERTICES1), C.ConvexHullVertices(EPA_VERTICES2)
    dist, p1, p2, simplex = gjk.gjk(c1, c2)
    np.testing.assert_allclose(p1, p2, atol=1e-6)
    mtv, faces, success = epa.epa(simplex, c1, c2)
    assert success and faces.shape[1:] == (4, 3)
    np.testing.assert_allclose(mtv, [-0.387287, 0.179576, -0.176204], atol=5e-7)
    # moving collider 2 by mtv (plus a little) separates the shapes
    c2b = C.ConvexHullVertices(EPA_VERTICES2 + mtv * (1.0 + 1e-3))
    assert gjk.gjk(c1, c2b)[0] > 0.0


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_undefined_simplices_are_reported_not_run(monkeypatch, kernel):
    """GJK may end with fewer than 4 simplex points; the reference's epa() then reads rows of an
    np.empty array (undefined).  With `n_points` those pairs get status 8 and the rest is
    unchanged."""
    import torch
    monkeypatch.setenv("D3D_EPA_KERNEL", kernel)
    rs = np.random.RandomState(5)
    cs = d3random.random_collider_set(rs, 800, center_scale=0.3)
    pairs = d3random.random_pairs(rs, len(cs), 6000)
    g = gjk.gjk_distance_batch(cs, pairs)
    hits = torch.nonzero(g.dist == 0.0).flatten()
    sub = pairs[hits.cpu().numpy()]
    npts = g.n_points[hits]
    assert int((npts < 4).sum()) > 0 and int((npts == 4).sum()) > 100
    res = epa.epa_batch(cs, sub, g.simplex[hits], n_points=npts).cpu()
    bad = npts.cpu().numpy() != 4
    assert np.all(res["status"][bad] == gjk.STATUS_EPA_BAD_SIMPLEX)
    assert not res["mtv"][bad].any() and not res["success"][bad].any()
    full = epa.epa_batch(cs, sub[~bad], g.simplex[hits][torch.from_numpy(~bad).cuda()]).cpu()
    assert np.array_equal(res["mtv"][~bad], full["mtv"]) and np.array_equal(res["status"][~bad], full["status"])


def test_degenerate_simplices_match_the_oracle(monkeypatch):
    """Simplices with a duplicated point or two points closer than the edge-matching epsilon
    (epa.py:189-191 matches loose edges by distance) exercise the coordinate-based edge matching
    where it differs from matching by vertex identity; results are the oracle's bit for bit."""
    import torch
    rs = np.random.RandomState(9)
    cs = d3random.random_collider_set(rs, 1200, names=d3random.PRIMITIVES + ("mesh",), center_scale=0.35,
                                      hull_vertices=(6, 40))
    pairs = d3random.random_pairs(rs, len(cs), 8000)
    g = gjk.gjk_distance_batch(cs, pairs)
    sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()
    sub = pairs[sel.cpu().numpy()]
    Y = g.simplex[sel].cpu().numpy().copy()
    Y[::7, 1] = Y[::7, 0]
    Y[3::11, 2] = Y[3::11, 3] + 1e-9
    res = epa.epa_batch(cs, sub, Y, want_faces=True).cpu()
    ref = O.epa(cs, sub, Y, return_faces=True, n_threads=O.max_threads())
    assert np.array_equal(res["status"], ref["status"])
    ok = ref["status"] != 7
    assert np.array_equal(res["mtv"][ok], ref["mtv"][ok])
    assert np.array_equal(res["n_faces"][ok], ref["n_faces"][ok]) and np.array_equal(res["iters"][ok], ref["iters"][ok])
    for q in np.where(ok)[0][:300]:
        n = ref["n_faces"][q]
        assert np.array_equal(res["faces"][q, :n], ref["faces"][q, :n])
    assert len(sub) > 1000 and ok.sum() > 500
    # thread-per-pair kernel: duplicated points share a vertex id, near-duplicates are kept as
    # "near pairs" that the edge matching treats as equal, like the reference's distance test
    monkeypatch.setenv("D3D_EPA_KERNEL", "thread")
    thr = epa.epa_batch(cs, sub, Y).cpu()
    assert np.array_equal(thr["status"], ref["status"])
    for key in ("mtv", "success", "n_faces", "iters"):
        assert np.array_equal(thr[key][ok], ref[key][ok]), key
    assert int(thr["deferred"][0]) < 0.15 * len(sub)   # near-pair list full or max_iter reached


def test_thread_and_warp_kernels_agree_on_a_large_mixed_batch(monkeypatch):
    import torch
    rs = np.random.RandomState(31)
    cs = d3random.random_collider_set(rs, 20000, names=d3random.PRIMITIVES + ("mesh", "cone"),
                                      center_scale=0.9, hull_vertices=(4, 60))
    pairs = d3random.random_pairs(rs, len(cs), 400000)
    g = gjk.gjk_distance_batch(cs, pairs)
    hits = torch.nonzero(g.dist == 0.0).flatten()
    sub, Y, npts = pairs[hits.cpu().numpy()], g.simplex[hits], g.n_points[hits]
    assert len(sub) > 50000
    thr = epa.epa_batch(cs, sub, Y, n_points=npts).cpu()
    monkeypatch.setenv("D3D_EPA_KERNEL", "warp")
    wrp = epa.epa_batch(cs, sub, Y, n_points=npts).cpu()
    assert int(wrp["deferred"][0]) == 0 and int(thr["deferred"][0]) < 0.6 * len(sub)   # hulls of 33..60 vertices
    assert np.array_equal(thr["status"], wrp["status"])
    ok = wrp["status"] != 7
    for key in ("mtv", "success", "n_faces", "iters"):
        assert np.array_equal(thr[key][ok], wrp[key][ok]), key
    assert np.array_equal(thr["iters"], wrp["iters"])


@pytest.mark.parametrize("kernel", ["warp", "thread"])
def test_empty_and_single_pair_batches(monkeypatch, kernel):
    monkeypatch.setenv("D3D_EPA_KERNEL", kernel)
    rs = np.random.RandomState(2)
    cs = d3random.random_collider_set(rs, 50, center_scale=0.2)
    pairs = d3random.random_pairs(rs, len(cs), 200)
    g = gjk.gjk_distance_batch(cs, pairs).cpu()
    sel = np.where((g["dist"] == 0.0) & (g["n_points"] == 4))[0]
    assert len(sel) > 3
    empty = epa.epa_batch(cs, pairs[:0], g["simplex"][:0]).cpu()
    assert empty["mtv"].shape == (0, 3) and empty["status"].shape == (0,)
    one = epa.epa_batch(cs, pairs[sel[:1]], g["simplex"][sel[:1]]).cpu()
    ref = O.epa(cs, pairs[sel[:1]], g["simplex"][sel[:1]])
    assert np.array_equal(one["status"], ref["status"])
    if ref["status"][0] != 7:
        assert np.array_equal(one["mtv"], ref["mtv"]) and np.array_equal(one["iters"], ref["iters"])
