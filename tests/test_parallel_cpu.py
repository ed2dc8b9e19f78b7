"""CPU: the multi-rank host logic on a world_size-2 gloo group (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distance3d_b200 import parallel


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for ws in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n, r, ws) for r in range(ws)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
                assert b == c
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
            assert parallel.shard_counts(n, ws) == sizes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert parallel.world() == (rank, world_size)
        # each rank owns a shard of a global pair list and produces a variable-length result
        n = 1001
        begin, end = parallel.shard_range(n)
        local = torch.arange(begin, end, dtype=torch.int64)
        hits = local[local % (3 + rank) == 0]          # data-dependent length
        pairs = torch.stack([hits, hits * 2], dim=1).to(torch.int32)
        full, counts = parallel.all_gather_varlen(pairs)
        # empty contribution from one rank must work too
        empty = pairs[:0] if rank == 1 else pairs
        full2, counts2 = parallel.all_gather_varlen(empty)
        np.save(os.path.join(out_dir, "full_%d.npy" % rank), full.numpy())
        np.save(os.path.join(out_dir, "counts_%d.npy" % rank), np.array(counts + counts2))
        np.save(os.path.join(out_dir, "full2_%d.npy" % rank), full2.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_all_gather_varlen_world_size_2(tmp_path):
    ws = 2
    mp.spawn(_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    expect = []
    expect2 = []
    for rank in range(ws):
        begin, end = parallel.shard_range(1001, rank, ws)
        local = np.arange(begin, end)
        hits = local[local % (3 + rank) == 0]
        expect.append(np.stack([hits, hits * 2], axis=1))
        expect2.append(expect[-1][:0] if rank == 1 else expect[-1])
    expect = np.concatenate(expect)
    expect2 = np.concatenate(expect2)
    for rank in range(ws):
        np.testing.assert_array_equal(np.load(tmp_path / ("full_%d.npy" % rank)), expect)
        np.testing.assert_array_equal(np.load(tmp_path / ("full2_%d.npy" % rank)), expect2)
    c0 = np.load(tmp_path / "counts_0.npy")
    assert c0[0] + c0[1] == len(expect) and c0[3] == 0


def test_peer_pair_buffer_contract_documented():
    """The fused self query + all-gather needs >= 2 GPUs of one NVLink domain; its check lives in
    scripts/check_fused_gather.py (run under torchrun on the GPU box).  Here: the C ABI exports it."""
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "distance3d_b200", "libd3d_b200.so")
    lib = ctypes.CDLL(so)
    assert hasattr(lib, "d3d_bvh_overlap_self_gather")
    lib.d3d_last_error_string.restype = ctypes.c_char_p
    assert lib.d3d_bvh_overlap_self_gather(None, ctypes.c_int64(0), 0, 1, None, ctypes.c_int64(0), None, None, None, None) == -1
    assert b"null argument" in lib.d3d_last_error_string()
