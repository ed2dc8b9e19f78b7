"""torchrun --nproc-per-node N scripts/check_fused_gather.py [n_boxes]
Self query of a replicated LBVH with the all-gather fused into the traversal kernel (peer stores
over NVLink, parallel.PeerPairBuffer) against (i) the single-GPU pair set and (ii) the
traversal + NCCL exact-size gather, with timings (max over ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from distance3d_b200 import _lib, aabb_tree, parallel, random as R

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for name, scale in (("dense", 2.0), ("constant_density", 2.0 * (n / 2000.0) ** (1.0 / 3.0))):
    dc = R.random_capsules_device(32, n, center_scale=scale, device=dev)
    aabb = _lib.aabb_device(dc)
    bvh = aabb_tree.Lbvh(aabb)
    mine, count = bvh.overlap_unique(rank, world)
    cmax = torch.tensor([count], dtype=torch.int64, device=dev); dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
    total = torch.tensor([count], dtype=torch.int64, device=dev); dist.all_reduce(total)
    buf = parallel.PeerPairBuffer(int(cmax.item()) + 1024)
    parallel.overlap_unique_fused_gather(bvh, buf)
    segs, counts = buf.segments()
    assert sum(counts) == int(total.item()), (counts, int(total.item()))
    key = lambda p: torch.minimum(p[:, 0], p[:, 1]).long() * n + torch.maximum(p[:, 0], p[:, 1]).long()
    fused = torch.sort(torch.cat([key(s) for s in segs]))[0]
    full, _ = bvh.overlap_unique(0, 1)
    ref = torch.sort(key(full))[0]
    assert torch.equal(fused, ref), "fused gather != single-GPU pair set"
    del full, ref, fused
    out = torch.empty((count, 2), dtype=torch.int32, device=dev)
    gathered = torch.empty((int(total.item()), 2), dtype=torch.int32, device=dev)
    t_query = timed(lambda: bvh.overlap_unique_async(out, rank, world))
    t_nccl = timed(lambda: (bvh.overlap_unique_async(out, rank, world), parallel.all_gather_varlen(out[:count], out=gathered)))
    t_fused = timed(lambda: parallel.overlap_unique_fused_gather(bvh, buf))
    recv = (int(total.item()) - count) * 8.0
    if rank == 0:
        print("%s n=%d world=%d: %d unique pairs; query only %.3f ms; query + NCCL exact-size gather %.3f ms; "
              "fused (peer stores in the traversal kernel) %.3f ms = %.3e pairs/s, %.0f GB/s received per GPU"
              % (name, n, world, int(total.item()), t_query, t_nccl, t_fused, int(total.item()) / t_fused * 1e3,
                 recv / (t_fused * 1e-3) / 1e9), flush=True)
    del buf, out, gathered, bvh
    torch.cuda.empty_cache()
dist.barrier(); dist.destroy_process_group()
