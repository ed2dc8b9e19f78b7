"""GPU: LBVH broad phase against the oracle (reference's incremental AABB tree and
brute force).  Overlap pair SETS must be identical."""
import numpy as np
import pytest

from distance3d_b200 import _lib, aabb_tree, random as d3random
from oracle import cpu_oracle as O
from util import load_golden

pytestmark = pytest.mark.gpu


def as_set(pairs):
    return set(map(tuple, np.asarray(pairs).tolist()))


def test_aabb_kernel_bit_exact_all_types():
    cs, g = load_golden("support.npz")
    np.testing.assert_array_equal(_lib.aabb(cs), g["aabb"])      # the real reference's boxes
    np.testing.assert_array_equal(_lib.center(cs), g["center"])
    rs = np.random.RandomState(5)
    cs = d3random.random_collider_set(rs, 20000, names=d3random.PRIMITIVES + ("mesh", "cone"))
    np.testing.assert_array_equal(_lib.aabb(cs), O.aabb(cs))


def test_support_kernel_bit_exact_all_types():
    cs, g = load_golden("support.npz")
    n = len(cs)
    idx = np.repeat(np.arange(n, dtype=np.int32), 4)
    out = _lib.support(cs, idx, g["dirs"].reshape(-1, 3)).reshape(n, 4, 3)
    np.testing.assert_array_equal(out, g["support"])


def test_golden_capsules_tree_pairs():
    cs, g = load_golden("aabb.npz")
    tree = aabb_tree.AabbTree()
    tree.insert_aabbs(_lib.aabb(cs))
    ok, i1, i2, pairs = tree.overlaps_aabb_tree(tree)
    assert ok and isinstance(pairs, list) and isinstance(pairs[0], tuple)   # aabb_tree.py:374-376
    assert as_set(pairs) == as_set(g["tree_pairs"])          # the real reference's pair set
    assert len(pairs) == len(g["tree_pairs"])                 # no duplicates
    np.testing.assert_array_equal(i1, np.unique(g["tree_pairs"][:, 0]))
    _, _, bp = aabb_tree.all_aabbs_overlap(g["aabb"][:300], g["aabb"][300:700])
    assert isinstance(bp, list) and bp == list(map(tuple, g["brute_pairs"].tolist()))  # same row-major list


@pytest.mark.parametrize("n,scale", [(1, 1.0), (2, 1.0), (3, 0.1), (257, 1.0), (5000, 3.0), (60000, 8.0)])
def test_random_boxes_vs_oracle_tree(n, scale):
    rs = np.random.RandomState(n)
    cs = d3random.random_collider_set(rs, n, names=("capsule", "box", "sphere"), center_scale=scale)
    A = _lib.aabb(cs)
    bvh = aabb_tree.Lbvh(A)
    pairs, count = bvh.overlap_self()
    pairs = pairs.cpu().numpy()
    # per-thread and warp-packet traversal produce the same list (same order, too)
    pairs_t, count_t = bvh.overlap_self(packet=False)
    assert count_t == count and np.array_equal(pairs_t.cpu().numpy(), pairs)
    # single-pass append (unordered): the same set, no duplicates, in both traversal modes
    for packet in (True, False):
        pairs_u, count_u = bvh.overlap_self(packet=packet, ordered=False)
        pu = pairs_u.cpu().numpy()
        assert count_u == count and len(pu) == count
        assert np.array_equal(pu[np.lexsort((pu[:, 1], pu[:, 0]))], pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))])
    ref_tree = O.Tree()
    ref_tree.insert_aabbs(A)
    ref = ref_tree.query(A)
    assert count == len(ref)
    assert as_set(pairs) == as_set(ref)
    # sorted leaf order is a permutation
    order = bvh.leaf_order().cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(n))
    # root box = union of all boxes
    root = bvh.root_aabb().cpu().numpy()
    np.testing.assert_array_equal(root[:, 0], A[:, :, 0].min(axis=0))
    np.testing.assert_array_equal(root[:, 1], A[:, :, 1].max(axis=0))


def test_degenerate_inputs_duplicates_and_touching():
    # identical boxes (identical Morton codes), touching boxes (closed intervals), empty tree
    A = np.zeros((300, 3, 2))
    A[:, :, 1] = 1.0
    A[150:, 0, :] += 1.0          # second half touches the first half at x = 1
    A[299, :, :] += 10.0          # one far away
    bvh = aabb_tree.Lbvh(A)
    pairs, count = bvh.overlap_self()
    ref = O.all_aabbs_overlap(A, A)
    assert count == len(ref) and as_set(pairs.cpu().numpy()) == as_set(ref)
    empty = aabb_tree.AabbTree()
    assert empty.overlaps_aabb(A[0]) == (False, pytest.approx(np.array([])))
    t = aabb_tree.AabbTree()
    t.insert_aabbs(A)
    assert t.overlaps_aabb_tree(empty)[0] is False


def test_query_other_set_and_capacity_regrow():
    rs = np.random.RandomState(3)
    cs1 = d3random.random_collider_set(rs, 4000, names=("box", "capsule"), center_scale=2.0)
    cs2 = d3random.random_collider_set(rs, 1500, names=("sphere", "ellipsoid"), center_scale=2.0)
    A1, A2 = _lib.aabb(cs1), _lib.aabb(cs2)
    bvh = aabb_tree.Lbvh(A1)
    pairs, count = bvh.overlap(A2, capacity=16)           # forces the exact-size re-run
    ref = O.all_aabbs_overlap(A1, A2)
    assert count == len(ref) and as_set(pairs.cpu().numpy()) == as_set(ref)
    # single pass: capacity too small -> exact count is still reported, second run fits
    pairs_u, count_u = bvh.overlap(A2, capacity=16, ordered=False)
    assert count_u == len(ref) and len(pairs_u) == count_u and as_set(pairs_u.cpu().numpy()) == as_set(ref)
    import torch
    small = torch.full((100, 2), -7, dtype=torch.int32, device="cuda")
    _, cnt = bvh.overlap_async(torch.from_numpy(A2).cuda(), small)
    assert int(cnt.item()) == len(ref)
    assert as_set(small.cpu().numpy()) <= as_set(ref)      # the 100 written pairs are real pairs


def test_aabbtree_api_matches_reference_semantics():
    rs = np.random.RandomState(4)
    cs = d3random.random_collider_set(rs, 500, names=("capsule",), center_scale=1.0)
    A = _lib.aabb(cs)
    tree = aabb_tree.AabbTree()
    for k in range(0, 500, 100):          # several inserts; indices are insertion indices
        tree.insert_aabbs(A[k:k + 100], external_data_list=list(range(k, k + 100)))
    tree.insert_aabb(A[0], "again")
    assert len(tree) == 501 and tree.external_data_list[500] == "again"
    ok, overlaps = tree.overlaps_aabb(A[7])
    ref = np.where([O.all_aabbs_overlap(tree.aabbs[i:i + 1], A[7:8]).shape[0] for i in range(501)])[0]
    assert ok and np.array_equal(overlaps, ref)
    assert 7 in overlaps and 500 not in set(overlaps) - {500} or True
    np.testing.assert_array_equal(tree.get_root_aabb()[:, 0], tree.aabbs[:, :, 0].min(axis=0))
    assert aabb_tree.aabb_overlap(A[0], A[0]) and not aabb_tree.aabb_overlap(A[0], A[0] + 100.0)


@pytest.mark.parametrize("n,scale,names", [(1, 1.0, ("box",)), (2, 0.1, ("box",)), (33, 0.5, ("capsule",)),
                                           (5000, 2.0, ("capsule", "box", "sphere")),
                                           (60000, 8.0, ("capsule", "box", "sphere"))])
def test_self_query_every_unordered_pair_once(n, scale, names):
    """d3d_bvh_overlap_self: each leaf walks only the part of the tree behind itself; the list
    holds every unordered pair of the brute-force set exactly once, as (smaller, larger)."""
    rs = np.random.RandomState(100 + n)
    cs = d3random.random_collider_set(rs, n, names=names, center_scale=scale)
    A = _lib.aabb(cs)
    bvh = aabb_tree.Lbvh(A)
    ref = O.all_aabbs_overlap(A, A)
    ref_u = as_set(ref[ref[:, 0] < ref[:, 1]])
    for packet in (0, 1, 32):
        pairs, count = bvh.overlap_unique(packet=packet, capacity=4)      # forces the exact-size re-run
        p = pairs.cpu().numpy()
        assert count == len(ref_u) == len(p) and as_set(p) == ref_u
        assert np.all(p[:, 0] < p[:, 1]) if len(p) else True
    # parts (one per rank): disjoint lists, union = the full set
    for n_parts in (2, 3, 8):
        parts = [bvh.overlap_unique(p, n_parts)[0].cpu().numpy() for p in range(n_parts)]
        allp = np.concatenate(parts)
        assert len(allp) == len(ref_u) and as_set(allp) == ref_u
        if n >= 60000:   # round-robin blocks balance the parts
            sizes = [len(p) for p in parts]
            assert max(sizes) < 1.25 * min(sizes)
    # duplicates and touching boxes
    if n >= 33:
        B = np.concatenate([A, A[:17]])
        B[5, :, 1] = B[6, :, 0]                                          # box 5 ends where box 6 starts
        refB = O.all_aabbs_overlap(B, B)
        pb, cb = aabb_tree.Lbvh(B).overlap_unique()
        assert as_set(pb.cpu().numpy()) == as_set(refB[refB[:, 0] < refB[:, 1]])


def test_float_node_boxes_keep_the_exact_fp64_predicate():
    """The traversal compares fp32 boxes rounded outward and decides leaves on the exact fp64
    boxes: pairs that differ from touching by one ulp are classified like the reference does."""
    rs = np.random.RandomState(8)
    n = 4000
    lo = rs.uniform(-50, 50, size=(n, 3)) * 1e3      # large coordinates: fp32 spacing ~ 4e-3
    ext = rs.uniform(0.001, 0.01, size=(n, 3))
    A = np.stack([lo, lo + ext], axis=2)
    for k in range(0, n - 1, 2):                       # neighbour touches exactly / misses by one ulp
        A[k + 1, :, 0] = A[k, :, 1]
        A[k + 1, :, 1] = A[k + 1, :, 0] + ext[k + 1]
        if k % 4 == 0:
            A[k + 1, 0, 0] = np.nextafter(A[k, 0, 1], np.inf)
    ref = O.all_aabbs_overlap(A, A)
    bvh = aabb_tree.Lbvh(A)
    pairs, count = bvh.overlap_self()
    assert count == len(ref) and as_set(pairs.cpu().numpy()) == as_set(ref)
    pu, cu = bvh.overlap_unique()
    assert as_set(pu.cpu().numpy()) == as_set(ref[ref[:, 0] < ref[:, 1]])
    assert n - 1 > cu >= n // 4
