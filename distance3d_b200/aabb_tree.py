"""AABB tree for broad-phase collision detection, on the GPU.

Drop-in for distance3d/aabb_tree.py: `AabbTree` (:15-191), `all_aabbs_overlap`
(:465-500) and `aabb_overlap` (:503-527).  The reference grows an incremental
binary tree on the host; here every `insert_aabbs` rebuilds a linear BVH on the
device (Morton codes, radix sort, Karras hierarchy, bottom-up refit; see
csrc/lbvh.cu) and queries run a stackless traversal.  Only overlap SETS are
contractual: indices returned by the queries are insertion indices
(0 .. n-1 in the order the boxes were inserted) and index `external_data_list`.

`Lbvh` is the batched device-level interface used by the pipeline.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_i64, c_size, ptr

INDEX_NONE = -1


class Lbvh:
    """Linear BVH over device-resident boxes ``aabbs f64[n,3,2]`` (torch tensor)."""

    def __init__(self, aabbs):
        torch = _lib.torch_cuda()
        if not isinstance(aabbs, torch.Tensor):
            aabbs = torch.from_numpy(np.ascontiguousarray(aabbs, dtype=np.float64)).to(
                torch.device("cuda", torch.cuda.current_device()))
        self.aabbs = aabbs.reshape(-1, 3, 2).contiguous()
        self.n = int(self.aabbs.shape[0])
        self.device = self.aabbs.device
        L = _lib.lib()
        nbytes = L.d3d_bvh_workspace_bytes(c_i64(self.n))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._visits = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._qws = None
        _lib._check(L.d3d_bvh_build(ptr(self.aabbs), c_i64(self.n), ptr(self.workspace),
                                    c_size(nbytes), _lib.stream_ptr()))
        self._order = None

    def leaf_order(self):
        """Object indices in Morton order (int32[n])."""
        if self._order is None:
            torch = _lib.torch_cuda()
            self._order = torch.empty(self.n, dtype=torch.int32, device=self.device)
            _lib._check(_lib.lib().d3d_bvh_leaf_order(ptr(self.workspace), c_i64(self.n),
                                                      ptr(self._order), _lib.stream_ptr()))
        return self._order

    def root_aabb(self):
        torch = _lib.torch_cuda()
        out = torch.empty((3, 2), dtype=torch.float64, device=self.device)
        _lib._check(_lib.lib().d3d_bvh_root_aabb(ptr(self.workspace), c_i64(self.n), ptr(out),
                                                 _lib.stream_ptr()))
        return out

    def overlap(self, query, capacity=None, order=None, out=None, count_visits=False, packet=None,
                ordered=True):
        """All (tree index, query index) pairs with overlapping boxes.

        Returns ``(pairs int32[count, 2], count)`` as device tensor / int.  ordered=True
        (default): two passes on the device, the exact count is known before the pair buffer
        is allocated (pass `out` to re-use a buffer; it is replaced when too small) and the
        pair order is reproducible.  ordered=False: ONE traversal that appends into a buffer
        of `capacity` pairs (default 16 per query, or the given `out`) and is repeated with the
        exact size if that was too small; the same set in an arbitrary order, about twice as
        fast.  `packet`: warp-packet traversal (default: on when the queries are spatially
        ordered, i.e. `order` is given).
        """
        if packet is None:
            packet = order is not None
        torch = _lib.torch_cuda()
        if not isinstance(query, torch.Tensor):
            query = torch.from_numpy(np.ascontiguousarray(query, dtype=np.float64)).to(self.device)
        query = query.reshape(-1, 3, 2).contiguous()
        # with `order` only the listed query boxes are processed (e.g. one rank's shard)
        nq = int(order.shape[0]) if order is not None else int(query.shape[0])
        L = _lib.lib()
        if not ordered:
            width = int(packet)  # True / False or an explicit packet width 2, 4, 8, 16, 32
            if out is None:
                out = torch.empty((max(int(capacity or 16 * nq), 1), 2), dtype=torch.int32,
                                  device=self.device)
            for _ in range(2):
                _lib._check(L.d3d_bvh_overlap(
                    ptr(self.workspace), c_i64(self.n), ptr(query), ptr(order), c_i64(nq),
                    ctypes.c_int(width), ptr(out), c_i64(out.shape[0]), ptr(self._count),
                    ptr(self._visits) if count_visits else None, None, c_size(0), _lib.stream_ptr()))
                count = int(self._count.item())
                if count <= out.shape[0]:
                    break
                out = torch.empty((count, 2), dtype=torch.int32, device=self.device)
            return out[:count], count
        qbytes = L.d3d_bvh_query_workspace_bytes(c_i64(nq))
        if self._qws is None or self._qws.numel() < qbytes:
            self._qws = torch.empty(qbytes, dtype=torch.uint8, device=self.device)
        _lib._check(L.d3d_bvh_overlap_count(
            ptr(self.workspace), c_i64(self.n), ptr(query), ptr(order), c_i64(nq),
            ctypes.c_int(1 if packet else 0), ptr(self._count),
            ptr(self._visits) if count_visits else None, ptr(self._qws), c_size(self._qws.numel()),
            _lib.stream_ptr()))
        count = int(self._count.item())
        if out is None or out.shape[0] < count:
            out = torch.empty((max(count, 1), 2), dtype=torch.int32, device=self.device)
        _lib._check(L.d3d_bvh_overlap_fill(
            ptr(self.workspace), c_i64(self.n), ptr(query), ptr(order), c_i64(nq),
            ctypes.c_int(1 if packet else 0), ptr(out), c_i64(out.shape[0]), ptr(self._qws),
            _lib.stream_ptr()))
        return out[:count], count

    def overlap_async(self, query, out, order=None, packet=None):
        """Single-pass traversal that appends into a caller-provided buffer without
        synchronising the host (arbitrary pair order).

        Returns ``(out, count)`` where `count` is a device int64 tensor with the exact number of
        pairs (entries beyond ``out.shape[0]`` are dropped; check `count` when in doubt)."""
        torch = _lib.torch_cuda()
        query = query.reshape(-1, 3, 2)
        nq = int(order.shape[0]) if order is not None else int(query.shape[0])
        if packet is None:
            packet = order is not None
        L = _lib.lib()
        _lib._check(L.d3d_bvh_overlap(
            ptr(self.workspace), c_i64(self.n), ptr(query), ptr(order), c_i64(nq),
            ctypes.c_int(int(packet)), ptr(out), c_i64(out.shape[0]), ptr(self._count), None,
            None, c_size(0), _lib.stream_ptr()))
        return out, self._count

    def overlap_unique_async(self, out, part=0, n_parts=1, packet=True, count_visits=False):
        """The tree against its own leaves: every unordered overlapping pair ONCE as (smaller,
        larger) object index, no (i, i); one traversal in which a leaf only walks the part of
        the tree behind itself (half the nodes and half the output of `overlap_self`).
        `part` of `n_parts` (one per rank, all holding the same tree): the leaves are dealt in
        blocks of 128 of the Morton order, round-robin; the parts' lists are disjoint and their
        union is the full set.  Appends into `out` int32[cap, 2] without synchronising the
        host; returns ``(out, count)`` with the exact count as a device int64 tensor (pairs
        beyond cap are dropped)."""
        _lib._check(_lib.lib().d3d_bvh_overlap_self(
            ptr(self.workspace), c_i64(self.n), ctypes.c_int(part), ctypes.c_int(n_parts),
            ctypes.c_int(int(packet)), ptr(out), c_i64(out.shape[0]), ptr(self._count),
            ptr(self._visits) if count_visits else None, _lib.stream_ptr()))
        return out, self._count

    def overlap_unique(self, part=0, n_parts=1, capacity=None, out=None, packet=True,
                       count_visits=False):
        """Synchronous form of :meth:`overlap_unique_async`: ``(pairs int32[count, 2], count)``;
        the traversal is repeated with the exact size when the buffer was too small."""
        torch = _lib.torch_cuda()
        if out is None:
            out = torch.empty((max(int(capacity or 8 * (self.n // n_parts + 1)), 1), 2),
                              dtype=torch.int32, device=self.device)
        for _ in range(2):
            self.overlap_unique_async(out, part, n_parts, packet, count_visits)
            count = int(self._count.item())
            if count <= out.shape[0]:
                break
            out = torch.empty((count, 2), dtype=torch.int32, device=self.device)
        return out[:count], count

    def overlap_self(self, capacity=None, out=None, count_visits=False, packet=None, ordered=True):
        """Tree against its own leaves, queries walked in Morton order."""
        return self.overlap(self.aabbs, capacity=capacity, order=self.leaf_order(), out=out,
                            count_visits=count_visits, packet=packet, ordered=ordered)

    def visits(self):
        """Node records fetched by the last overlap(count_visits=True) call."""
        return int(self._visits.item())

    def rebuild(self, aabbs=None):
        """Rebuild the tree in place (same workspace) after the boxes changed."""
        if aabbs is not None:
            self.aabbs.copy_(aabbs.reshape(-1, 3, 2))
        _lib._check(_lib.lib().d3d_bvh_build(ptr(self.aabbs), c_i64(self.n), ptr(self.workspace),
                                             c_size(self.workspace.numel()), _lib.stream_ptr()))
        self._order = None


def brute_force_pairs(aabbs1, aabbs2, capacity=None):
    """Device brute force: (pairs int32[count,2] sorted row-major, count)."""
    torch = _lib.torch_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())

    def dev_boxes(a):
        if isinstance(a, torch.Tensor):
            return a.to(dev).reshape(-1, 3, 2).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev).reshape(-1, 3, 2)

    a1, a2 = dev_boxes(aabbs1), dev_boxes(aabbs2)
    n1, n2 = int(a1.shape[0]), int(a2.shape[0])
    count_t = torch.zeros(1, dtype=torch.int64, device=dev)
    capacity = capacity or max(1024, 8 * (n1 + n2))
    L = _lib.lib()
    while True:
        out = torch.empty((capacity, 2), dtype=torch.int32, device=dev)
        _lib._check(L.d3d_aabb_overlap_brute(ptr(a1), c_i64(n1), ptr(a2), c_i64(n2), ptr(out),
                                             c_i64(capacity), ptr(count_t), _lib.stream_ptr()))
        count = int(count_t.item())
        if count <= capacity:
            break
        capacity = count
    pairs = out[:count]
    if count:
        key = pairs[:, 0].to(torch.int64) * max(n2, 1) + pairs[:, 1].to(torch.int64)
        pairs = pairs[torch.argsort(key)]
    return pairs, count


def all_aabbs_overlap(aabbs1, aabbs2):
    """Brute-force overlap of two box lists (reference: aabb_tree.py:465-500).

    Returns ``(unique indices 1, unique indices 2, pairs)``; pairs is the reference's list
    of ``(i, j)`` tuples in row-major order (`brute_force_pairs` keeps them on the device).
    """
    pairs, _ = brute_force_pairs(aabbs1, aabbs2)
    pairs = pairs.cpu().numpy().astype(np.int64)
    return np.unique(pairs[:, 0]), np.unique(pairs[:, 1]), list(map(tuple, pairs.tolist()))


def aabb_overlap(aabb1, aabb2):
    """Closed-interval overlap of two boxes (reference: aabb_tree.py:503-527): six
    comparisons of host scalars, evaluated where the two boxes are (the batched forms are
    `brute_force_pairs` and `Lbvh.overlap`)."""
    a = np.asarray(aabb1, dtype=np.float64)
    b = np.asarray(aabb2, dtype=np.float64)
    return bool(np.all(a[:, 0] <= b[:, 1]) and np.all(a[:, 1] >= b[:, 0]))


class AabbTree:
    """AABB tree with the reference's interface (aabb_tree.py:15-191)."""

    def __init__(self):
        self.root = INDEX_NONE
        self.filled_len = 0
        self.aabbs = np.empty((0, 3, 2))
        self.external_data_list = []
        self.insert_index_list = []
        self.insert_index_max = 0
        self._bvh = None

    def __len__(self):
        return len(self.aabbs)

    def insert_aabbs(self, aabbs, external_data_list=None, pre_insertion_methode="none"):
        """Insert boxes (reference: aabb_tree.py:31-101); the device tree is rebuilt lazily.

        `pre_insertion_methode` only influenced the shape of the reference's
        insertion tree and is accepted for compatibility.
        """
        aabbs = np.asarray(aabbs, dtype=np.float64).reshape(-1, 3, 2)
        n = len(aabbs)
        if n == 0:
            return
        assert external_data_list is None or len(external_data_list) == n
        self.aabbs = np.concatenate((self.aabbs, aabbs), axis=0)
        self.external_data_list += list(external_data_list) if external_data_list is not None \
            else [None] * n
        self.insert_index_list.extend(range(self.insert_index_max, self.insert_index_max + n))
        self.insert_index_max += n
        self.filled_len = len(self.aabbs)
        self.root = 0
        self._bvh = None

    def insert_aabb(self, aabb, external_data=None):
        """Insert a single box (reference: aabb_tree.py:103-115)."""
        self.insert_aabbs([aabb], [external_data], pre_insertion_methode="none")

    def _tree(self):
        if self._bvh is None:
            self._bvh = Lbvh(self.aabbs)
        return self._bvh

    def overlaps_aabb_tree(self, other):
        """Overlaps with another tree (reference: aabb_tree.py:121-159).

        Returns ``(is_overlapping, unique indices of self, unique indices of other,
        pairs)``; pairs is the reference's list of ``(index in self, index in other)``
        tuples, here sorted by (other, self) - the reference's order is that of its tree
        traversal and not part of the contract.  `Lbvh.overlap` keeps the pairs on the device.
        """
        if len(self) == 0 or len(other) == 0:
            return False, np.array([]), np.array([]), []
        bvh = self._tree()
        if other is self:
            pairs, _ = bvh.overlap_self()
        else:
            pairs, _ = bvh.overlap(other.aabbs)
        pairs = pairs.cpu().numpy().astype(np.int64)
        pairs = pairs[np.lexsort((pairs[:, 0], pairs[:, 1]))]
        return (len(pairs) > 0, np.unique(pairs[:, 0]), np.unique(pairs[:, 1]),
                list(map(tuple, pairs.tolist())))

    def overlaps_aabb(self, aabb):
        """Leaves overlapping one box (reference: aabb_tree.py:161-181)."""
        if len(self) == 0:
            return False, np.array([])
        pairs, _ = self._tree().overlap(np.asarray(aabb, dtype=np.float64).reshape(1, 3, 2))
        overlaps = np.sort(pairs[:, 0].cpu().numpy().astype(np.int64))
        return len(overlaps) > 0, overlaps

    def get_root_aabb(self):
        """AABB of the whole tree (reference: aabb_tree.py:183-191)."""
        return self._tree().root_aabb().cpu().numpy()
