"""torchrun --nproc-per-node 2 scripts/check_sharded_pipeline.py
Every rank runs the pipeline on its shard of the queries (replicated BVH); the union of
the ranks' candidate / contact lists must equal the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from distance3d_b200 import pipeline, parallel, random as R

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rs = np.random.RandomState(5)
cs = R.random_collider_set(rs, 200000, names=R.PRIMITIVES + ("mesh",), center_scale=18.0)
sharded = pipeline.collide(cs, shard=True)
cand, counts = parallel.all_gather_varlen(sharded.candidates)
dists, _ = parallel.all_gather_varlen(sharded.gjk.dist)
full = pipeline.collide(cs, shard=False)
key = lambda p: p[:, 0].long() * 2**31 + p[:, 1].long()
k_s, o_s = torch.sort(key(cand)); k_f, o_f = torch.sort(key(full.candidates))
assert torch.equal(k_s, k_f), "candidate sets differ"
assert torch.equal(dists[o_s], full.gjk.dist[o_f]), "distances differ"
n_hits = torch.tensor([sharded.hits.numel()], device="cuda"); dist.all_reduce(n_hits)
assert int(n_hits.item()) == full.hits.numel()
if rank == 0:
    print("sharded pipeline == single GPU: %d candidates (per rank %s), %d contacts" % (len(k_f), counts, full.hits.numel()))
dist.barrier(); dist.destroy_process_group()
