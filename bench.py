#!/usr/bin/env python
"""Benchmark of the hot path: batched GJK distance on random primitive pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[0] tiled, SURVEY.md section 8d "C1"): every pair
owns two fresh random colliders drawn uniformly from {sphere, ellipsoid, capsule,
cylinder, box} with the default scales of the reference's generators
(benchmarks/benchmark_gjk.py:7-20, distance3d/random.py:200-362); PAIRS pairs per
GPU (weak scaling: pairs are independent, every rank owns its shard, no
data-path collective).  A step = one d3d_gjk_distance pass over the rank's shard.

After the headline the same run measures BASELINE configs[1..4] at their stated sizes on all
N ranks, with the collectives the multi-GPU design needs inside the timed regions
(`broad_phase`, `epa`, `self_collision`, `pipeline` blocks of the JSON line):

  C2  1 M capsules: every rank builds the same LBVH (replica) and walks its 1/N of the leaves
      (strong scaling); the ranks' pair lists are all-gathered (exact-size slices)
  C3  4 Mi intersecting hull pairs (64-256 vertices) per GPU through EPA; mtv all-gathered
  C4  10 M joint configurations per GPU: FK + AABB + white-list filter + GJK; masks all-gathered
  C5  2 M mixed shapes per GPU (16 M on 8 GPUs): replicated LBVH over ALL shapes, this rank's
      leaves as queries, GJK on the candidates, EPA on the hits, contacts all-gathered

One JSON line is printed by rank 0 (see DESIGN.md "Measurement").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# stdout carries the ONE JSON line; whatever NCCL logs ("NCCL version ..." at NCCL_DEBUG=VERSION)
# goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

FLOP_PER_ITER = 350.0  # algorithmic flop per GJK iteration (primitives)
BYTES_PER_PAIR = 496.0  # 336 B in + 160 B out
NCU_PROFILE = os.path.join("profiles", "r02_ncu_k_gjk_thread_v8.txt")  # roofline.traffic is read from it
# reference's own numba path (gjk.gjk on 3000 primitive pairs, seed 84, JIT warm, best of 3), timed in
# the BUILD container (1 core of an 8-vCPU Xeon), not on the GPU box: its sources cannot travel
NUMBA_REFERENCE = {"value": 3.03e3, "unit": "pairs/s", "cores": 1, "kind": "reference (numba)",
                   "where": "build container, not the GPU box", "sample": "3000 pairs, seed 84"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=4 * 1024 * 1024, help="pairs per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2 * 1024 * 1024,
                    help="pairs timed on the host cores for cpu_baseline")
    ap.add_argument("--seed", type=int, default=84)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the broad-phase extra metrics")
    ap.add_argument("--capsules", type=int, default=1000000, help="C2: capsules in the broad phase")
    ap.add_argument("--only", default="", choices=["", "broad", "epa", "selfcollision", "pipeline", "libccd", "hydroelastic"],
                    help="run one secondary block only and print it (development)")
    ap.add_argument("--epa-pairs", type=int, default=4 * 1024 * 1024, help="C3: EPA pairs per GPU")
    ap.add_argument("--configurations", type=int, default=10000000, help="C4: joint configurations per GPU")
    ap.add_argument("--shapes", type=int, default=2000000, help="C5: shapes per GPU")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of the secondary blocks")
    return ap.parse_args()


def make_workload(seed, n_pairs):
    """2 * n_pairs random primitives; pair k = (2k, 2k + 1)."""
    from distance3d_b200 import random as d3random
    rs = np.random.RandomState(seed)
    cs = d3random.random_collider_set(rs, 2 * n_pairs, names=d3random.PRIMITIVES)
    pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(n_pairs, 2)
    return cs, pairs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_samples(self, n, timeout=15.0):
        """Block until nvidia-smi has delivered n more samples (it takes a second to start,
        longer when eight ranks start at once)."""
        if self.proc is None:
            return
        target = len(self.lines) + n
        t0 = time.time()
        while len(self.lines) < target and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measure_fp64_peak(torch, _lib):
    """Dependent-free FP64 FMA loop -> TFLOP/s (roofline denominator, measured live)."""
    L = _lib.lib()
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    sm = ctypes.c_int(0)
    L.d3d_device_info(ctypes.byref(sm), None, None)
    blocks, iters = sm.value * 8, 1 << 15
    s = _lib.stream_ptr()
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.d3d_fp64_peak_probe(_lib.ptr(scratch), ctypes.c_int(blocks), ctypes.c_int(iters), s)
        e1.record()
        torch.cuda.synchronize()
        flop = blocks * 256 * 8.0 * iters * 2.0
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def make_capsules(n, center_scale=2.0, seed=32):
    """BASELINE configs[1] (vis_capsules_benchmark.py:22-30 scaled up): n random capsules."""
    from distance3d_b200 import random as d3random, pack
    rs = np.random.RandomState(seed)
    pose = d3random.random_transforms(rs, n)
    pose[:, :3, 3] *= center_scale
    param = np.zeros((n, 3))
    param[:, 0] = (1.0 - rs.rand(n)) * 0.1
    param[:, 1] = (1.0 - rs.rand(n)) * 0.5
    z = np.zeros(n, dtype=np.int32)
    return pack.ColliderSet(np.full(n, pack.CAPSULE, dtype=np.int32), pose, param, z, z, np.zeros((0, 3)))


def timed_steps(torch, dist, world, dev, fn, steps, warmup):
    """max-over-ranks CUDA-event time per step of fn()."""
    for _ in range(max(warmup, 3)):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(torch, dist, world, dev, x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


def replicas_identical(torch, dist, world, t):
    """All ranks generated the same device-resident set (replicated BVH needs it)."""
    if world == 1:
        return True
    c = t.double().sum().reshape(1)
    lo, hi = c.clone(), c.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).item())


def pair_keys(torch, pairs, n):
    lo = torch.minimum(pairs[:, 0], pairs[:, 1]).long()
    hi = torch.maximum(pairs[:, 0], pairs[:, 1]).long()
    return lo * n + hi


NVLINK_PEAK_GBS = 770.0  # measured peer copy per direction on this pool (B200_PROFILING.md)


def block_broad_phase(args, torch, dist, rank, world, dev, hbm_peak):
    """C2: AABBs of n capsules; every rank builds the same LBVH and walks its share of the
    leaves (strong scaling); every unordered overlapping pair once; lists all-gathered."""
    from distance3d_b200 import _lib, aabb_tree, parallel, random as d3random
    n = args.capsules
    steps = args.extra_steps
    out = {"capsules": n, "scaling": "strong",
           "design": "replicated LBVH (built redundantly on every GPU), 128-leaf blocks dealt round-robin to the GPUs, "
                     "d3d_bvh_overlap_self: every unordered pair once; all-gather of exact-size slices"}
    for name, scale in (("dense", 2.0), ("constant_density", 2.0 * (n / 2000.0) ** (1.0 / 3.0))):
        dc = d3random.random_capsules_device(32, n, center_scale=scale, device=dev)
        aabb = _lib.aabb_device(dc)
        same = replicas_identical(torch, dist, world, aabb)
        bvh = aabb_tree.Lbvh(aabb)
        n_mine = len(range(rank, (n + 127) // 128, world)) * 128   # leaves of this rank's blocks
        pairs, count = bvh.overlap_unique(rank, world, count_visits=True)
        visits = bvh.visits()
        buf = torch.empty((max(count, 1), 2), dtype=torch.int32, device=dev)
        total = int(all_sum(torch, dist, world, dev, count))
        total_visits = int(all_sum(torch, dist, world, dev, visits))
        gathered = torch.empty((max(total, 1), 2), dtype=torch.int32, device=dev)
        build_ms = timed_steps(torch, dist, world, dev, lambda: bvh.rebuild(), steps, 3)
        query_ms = timed_steps(torch, dist, world, dev,
                               lambda: bvh.overlap_unique_async(buf, rank, world), steps, 3)
        res = {}

        def gather():
            res["g"] = parallel.all_gather_varlen(buf[:count], out=gathered)

        def whole():
            bvh.rebuild()
            bvh.overlap_unique_async(buf, rank, world)
            gather()
        gather_ms = timed_steps(torch, dist, world, dev, gather, steps, 3) if world > 1 else 0.0
        whole_ms = timed_steps(torch, dist, world, dev, whole, steps, 3)
        fused = None
        if world > 1:
            # the all-gather fused into the traversal kernel: peer stores over NVLink into a
            # symmetric buffer (parallel.PeerPairBuffer), no separate collective
            cmax = torch.tensor([count], dtype=torch.int64, device=dev)
            dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
            pbuf = parallel.PeerPairBuffer(int(cmax.item()) + 1024)
            parallel.overlap_unique_fused_gather(bvh, pbuf)
            segs, seg_counts = pbuf.segments()
            fused_ok = sum(seg_counts) == total and bool(
                torch.equal(torch.sort(pair_keys(torch, torch.cat(segs), n))[0],
                            torch.sort(pair_keys(torch, parallel.all_gather_varlen(buf[:count])[0], n))[0]))
            fused_ms = timed_steps(torch, dist, world, dev,
                                   lambda: parallel.overlap_unique_fused_gather(bvh, pbuf), steps, 3)
            fused = {"query_and_gather_ms": fused_ms, "overlap_pairs_per_s": total / (fused_ms * 1e-3),
                     "nvlink_gbs_received_per_gpu": (total - count) * 8.0 / (fused_ms * 1e-3) / 1e9,
                     "equals_nccl_gather": fused_ok,
                     "how": "d3d_bvh_overlap_self_gather: the traversal kernel stores every flushed chunk into all "
                            "GPUs' buffers through peer pointers; two stream barriers included"}
            del pbuf, segs
        recv_bytes = (total - count) * 8.0
        # the ordered form the reference returns (both orientations + (i, i)) for continuity with round 1
        full_buf = torch.empty((2 * total + n, 2), dtype=torch.int32, device=dev) if world == 1 else None
        ordered_ms = None
        if world == 1:
            ordered_ms = timed_steps(torch, dist, world, dev, lambda: bvh.overlap_async(
                bvh.aabbs, full_buf, order=bvh.leaf_order(), packet=True), 3, 2)
            del full_buf
        q_bytes = n_mine * 64.0 + count * 8.0 + visits * 32.0
        q_bytes_min = n_mine * 64.0 + count * 8.0 + min(visits * 32.0, n * 96.0)
        entry = {
            "center_scale": scale, "unique_overlap_pairs": total,
            "ordered_form_pairs": 2 * total + n, "replicas_identical": same,
            "node_visits": total_visits, "traversal": "self query, 8-wide packets",
            "build_ms": build_ms, "query_ms": query_ms, "gather_ms": gather_ms,
            "build_query_gather_ms": whole_ms,
            "build_aabbs_per_s": n / (build_ms * 1e-3),
            "overlap_pairs_per_s": total / (query_ms * 1e-3),
            "ordered_form_pairs_per_s": (2 * total + n) / (query_ms * 1e-3),
            "overlap_pairs_per_s_with_gather": total / ((query_ms + gather_ms) * 1e-3),
            "queries_per_s": n / (query_ms * 1e-3),
            "gather": {"bytes_received_per_gpu": recv_bytes,
                       "nvlink_gbs_per_gpu": recv_bytes / (gather_ms * 1e-3) / 1e9 if gather_ms else None,
                       "nvlink_peak_gbs": NVLINK_PEAK_GBS,
                       "frac": recv_bytes / (gather_ms * 1e-3) / 1e9 / NVLINK_PEAK_GBS if gather_ms else None},
            "roofline_query": {"bound": "hbm", "achieved": q_bytes / (query_ms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s",
                               "frac": q_bytes / (query_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes": "rank 0: Q*64 + pairs*8 + node_visits*32 (compact records); "
                                        "re-visits are served by L1 / L2",
                               "frac_compulsory": q_bytes_min / (query_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes_compulsory": "Q*64 + pairs*8 + min(node_visits*32, n*96)"},
            "roofline_build": {"bound": "hbm", "achieved": n * 176.0 / (build_ms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s",
                               "frac": n * 176.0 / (build_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes": "176 B per primitive (SURVEY 8d)"},
        }
        if ordered_ms is not None:
            entry["ordered_form_query_ms"] = ordered_ms   # d3d_bvh_overlap of all boxes (round-1 metric)
        if fused is not None:
            entry["fused_gather"] = fused
        if world > 1:   # the gathered list is the union of the ranks' disjoint lists
            keys = pair_keys(torch, res["g"][0], n)
            entry["gathered_pairs_unique"] = bool(torch.unique(keys).numel() == total)
            del keys
        else:
            # set parity at full size against an independent kernel: brute force over all n^2 boxes
            t0 = time.perf_counter()
            bp, bc = aabb_tree.brute_force_pairs(aabb, aabb, capacity=2 * total + n + 16)
            kb = torch.sort(pair_keys(torch, bp[bp[:, 0] < bp[:, 1]], n))[0]
            kl = torch.sort(pair_keys(torch, buf[:count], n))[0]
            entry["parity_vs_brute_force"] = {"boxes": n, "brute_pairs": int(bc),
                                              "sets_equal": bool(bc == 2 * total + n and torch.equal(kb, kl)),
                                              "seconds": time.perf_counter() - t0}
            del bp, kb, kl
        out[name] = entry
        del buf, gathered, bvh, pairs, res
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_cpu_baseline:
        # reference algorithm (incremental tree + per-box stack query) on a bounded sample
        from oracle import cpu_oracle
        m = min(n, 50000)
        A = cpu_oracle.aabb(make_capsules(m, 2.0 * (m / 2000.0) ** (1.0 / 3.0)))
        t0 = time.perf_counter()
        tree = cpu_oracle.Tree()
        tree.insert_aabbs(A)
        t1 = time.perf_counter()
        ref_pairs = tree.query(A)
        t2 = time.perf_counter()
        out["cpu_baseline"] = {
            "kind": "port", "cores": 1, "sample": "%d capsules at constant density" % m,
            "build_aabbs_per_s": m / (t1 - t0), "queries_per_s": m / (t2 - t1),
            "overlap_pairs_per_s": len(ref_pairs) / (t2 - t1)}
    return out


def block_epa(args, torch, dist, rank, world, dev):
    """C3: EPA on intersecting convex-hull pairs with 64-256 vertices each, args.epa_pairs per GPU."""
    from distance3d_b200 import gjk, epa, random as d3random
    want = args.epa_pairs
    n_gen = int(want * 1.06) + 1024          # ~97 % of the generated pairs intersect with a full simplex
    dc = d3random.random_collider_set_device(args.seed + 7 + 1000 * rank, 2 * n_gen, names=("mesh",),
                                             center_scale=0.5, hull_vertices=(64, 256), hull_library=4096,
                                             device=dev)
    pairs = torch.arange(2 * n_gen, dtype=torch.int32, device=dev).reshape(n_gen, 2)
    g = gjk.gjk_distance_batch(dc, pairs, want_points=False)
    ms_gjk = timed_steps(torch, dist, world, dev, lambda: gjk.gjk_distance_batch(dc, pairs, out=g), 2, 1)
    sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()[:want]
    n = int(sel.numel())
    pairs_d = pairs[sel].contiguous()
    Y = g.simplex[sel].contiguous()
    del g
    res = {}
    gathered = torch.empty((world * n, 3), dtype=torch.float64, device=dev) if world > 1 else None
    same_n = all_sum(torch, dist, world, dev, n) == world * n

    def step():
        res["r"] = epa.epa_batch(dc, pairs_d, Y)
        if world > 1 and same_n:
            dist.all_gather_into_tensor(gathered, res["r"].mtv)
    ms = timed_steps(torch, dist, world, dev, step, args.extra_steps, 3)
    total = all_sum(torch, dist, world, dev, n)
    r = res["r"]
    out = {"metric": "epa_pairs_per_s", "value": total / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
           "scaling": "weak",
           "config": {"workload": "C3: EPA on intersecting convex hulls, 64-256 vertices, 4-point GJK simplex, "
                                  "two fresh hulls per pair (4096 shapes, unique poses), center_scale 0.5",
                      "pairs_per_gpu": n, "hull_pairs_generated": n_gen,
                      "vertex_pool_gb": float(dc.verts.numel() * 8 / 1e9),
                      "collective": "all_gather_into_tensor of mtv inside the step" if world > 1 else "none"},
           "mean_epa_iterations": float(r.iters.double().mean().item()),
           "max_faces_assert_rate": float((r.status == 7).double().mean().item()),
           "converged_rate": float(r.success.double().mean().item()),
           "gjk_hull_pairs_per_s": world * n_gen / (ms_gjk * 1e-3)}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cpu_oracle
        m = min(n, 20000)
        threads = cpu_oracle.max_threads()
        sub_pairs = pairs_d[:m]
        idx = sub_pairs.reshape(-1)
        cs = d3random.device_set_to_host(dc, idx)
        host_pairs = np.arange(2 * m, dtype=np.int32).reshape(m, 2)
        t0 = time.perf_counter()
        ref = cpu_oracle.epa(cs, host_pairs, Y[:m].cpu().numpy(), n_threads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
                               "sample": "first %d EPA pairs of rank 0" % m}
        ok = ref["status"] != 7
        out["parity_on_cpu_sample"] = {
            "pairs": m, "bit_exact_mtv": bool(np.array_equal(r.mtv[:m].cpu().numpy()[ok], ref["mtv"][ok])),
            "status_equal": bool(np.array_equal(r.status[:m].cpu().numpy(), ref["status"]))}
    del dc, Y, pairs_d, res
    torch.cuda.empty_cache()
    return out


def block_self_collision(args, torch, dist, rank, world, dev):
    """C4: robot arm (6 revolute joints, 8 cylinders), q ~ U(-pi, pi)^6, args.configurations per GPU."""
    from distance3d_b200 import broad_phase, self_collision
    from distance3d_b200.urdf import UrdfTransformManager
    data = os.path.join(REPO, "tests", "data")
    tm = UrdfTransformManager()
    with open(os.path.join(data, "robot_arm.urdf")) as f:
        tm.load_urdf(f.read(), mesh_path=data)
    bvh = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
    bvh.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
    model = self_collision.RobotModel(tm, bvh)
    n = args.configurations
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed + 11 + 1000 * rank)
    q = (torch.rand((n, 6), generator=gen, device=dev, dtype=torch.float64) * 2.0 - 1.0) * np.pi
    res = {}
    gathered = torch.empty((world * n, model.n_frames), dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        res["r"] = model.detect_batch(q)
        if world > 1:
            dist.all_gather_into_tensor(gathered, res["r"][0])
    ms = timed_steps(torch, dist, world, dev, step, args.extra_steps, 3)
    mask, n_cand = res["r"]
    out = {"metric": "self_collision_configurations_per_s", "value": world * n / (ms * 1e-3),
           "unit": "configurations/s", "ms_per_step": ms, "scaling": "weak",
           "config": {"workload": "C4: URDF arm (6 joints, 8 cylinders, 17 candidate pairs), "
                                  "FK + AABB + white-list filter + GJK intersection",
                      "configurations_per_gpu": n,
                      "collective": "all_gather_into_tensor of the 8-bit masks inside the step" if world > 1 else "none"},
           "colliding_fraction": float((mask.sum(dim=1) > 0).double().mean().item()),
           "narrow_phase_candidates_per_configuration": n_cand / n}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cpu_oracle
        m = min(n, 200000)
        threads = cpu_oracle.max_threads()
        kin = tm.compile_kinematics(model.frames, "origin")
        qs = q[:m].cpu().numpy()
        t0 = time.perf_counter()
        ref_mask, _ = cpu_oracle.self_collision_masks(model.template, kin, model.pattern, qs, threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m / dt, "unit": "configurations/s", "cores": threads,
                               "kind": "port", "sample": "first %d configurations" % m}
        out["parity_on_cpu_sample"] = {
            "configurations": m,
            "masks_equal": bool(np.array_equal(mask[:m].cpu().numpy(), ref_mask))}
    del q, res, gathered
    torch.cuda.empty_cache()
    return out


def block_pipeline(args, torch, dist, rank, world, dev):
    """C5: args.shapes mixed shapes per GPU; LBVH over ALL shapes on every GPU (replica), this
    rank's leaves as queries, GJK on the candidates, EPA on the hits, contacts all-gathered."""
    from distance3d_b200 import pipeline, parallel, random as d3random
    n = args.shapes * world
    names = d3random.PRIMITIVES + ("mesh",)
    scale = 0.33 * n ** (1.0 / 3.0)   # ~ 10-30 AABB overlaps per shape
    dc = d3random.random_collider_set_device(args.seed + 13, n, names=names, center_scale=scale,
                                             hull_vertices=(10, 10), device=dev)
    same = replicas_identical(torch, dist, world, dc.pose)
    res = {}
    cap = int(12 * args.shapes)
    stage = {}

    def step():
        r = pipeline.collide(dc, shard=True, candidate_capacity=cap, timings=stage)
        contacts = r.candidates[r.hits]
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        res["contacts"], res["counts"] = parallel.all_gather_varlen(contacts)
        if r.epa is not None and world > 1:
            res["mtv"], _ = parallel.all_gather_varlen(r.epa.mtv)
        ev2 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        stage["gather"] = (ev, ev2)
        res["r"] = r
    ms = timed_steps(torch, dist, world, dev, step, args.extra_steps, 3)
    torch.cuda.synchronize()
    r = res["r"]
    cand_total = all_sum(torch, dist, world, dev, r.candidates.shape[0])
    hits_total = all_sum(torch, dist, world, dev, r.hits.numel())
    stages_ms = {k: float(a.elapsed_time(b)) for k, (a, b) in stage.items()}
    out = {"metric": "pipeline_shapes_per_s", "value": n / (ms * 1e-3), "unit": "shapes/s", "ms_per_step": ms,
           "scaling": "weak (queries per GPU fixed, the replicated tree grows with N)",
           "config": {"workload": "C5: mixed random shapes (5 primitives + 10-vertex hulls), LBVH build + self "
                                  "overlap query + GJK distance on candidates + EPA on hits + contact all-gather",
                      "shapes_total": n, "query_shapes_per_gpu": args.shapes, "center_scale": scale,
                      "replicas_identical": same},
           "aabb_overlaps_per_shape": 2.0 * cand_total / n,
           "candidate_pairs": int(cand_total),
           "candidate_pairs_per_s": cand_total / (ms * 1e-3),
           "contacts": int(hits_total),
           "gathered_contacts": int(res["contacts"].shape[0]),
           "stage_ms_rank0": stages_ms,
           "epa_share_of_step": stages_ms.get("epa", 0.0) / ms}
    if rank == 0 and not args.no_cpu_baseline:
        # reference algorithms on a bounded sample: incremental tree + stack queries, GJK, EPA
        from oracle import cpu_oracle
        m = min(n, 50000)
        threads = cpu_oracle.max_threads()
        sub_d = d3random.random_collider_set_device(args.seed + 17, m, names=names,
                                                    center_scale=0.33 * m ** (1.0 / 3.0),
                                                    hull_vertices=(10, 10), device=dev)
        sub = d3random.device_set_to_host(sub_d)
        t0 = time.perf_counter()
        A = cpu_oracle.aabb(sub)
        tree = cpu_oracle.Tree()
        tree.insert_aabbs(A)
        pr = tree.query(A)
        cand = pr[pr[:, 0] < pr[:, 1]]
        g = cpu_oracle.gjk_distance(sub, cand, n_threads=threads)
        selc = (g["dist"] == 0.0) & (g["n_points"] == 4)
        e = cpu_oracle.epa(sub, cand[selc], g["Y"][selc], n_threads=threads)
        dt = time.perf_counter() - t0
        if world > 1:
            gpu = None   # collide() would shard; the sample parity is checked in the N = 1 run
        else:
            gpu = pipeline.collide(sub_d, shard=False)
        out["cpu_baseline"] = {"value": m / dt, "unit": "shapes/s", "cores": threads, "kind": "port",
                               "sample": "%d shapes at the same density (tree build / query single-threaded, "
                                         "GJK / EPA on all threads)" % m}
        if gpu is not None:
            gc = gpu.candidates.cpu().numpy()
            ko = np.argsort(gc[:, 0].astype(np.int64) * m + gc[:, 1])
            kr = np.argsort(cand[:, 0].astype(np.int64) * m + cand[:, 1])
            sets_equal = len(gc) == len(cand) and bool(np.array_equal(gc[ko], cand[kr]))
            gd = gpu.gjk.dist.cpu().numpy()
            out["parity_on_cpu_sample"] = {
                "shapes": m, "candidate_sets_equal": sets_equal,
                "distances_bit_exact": bool(sets_equal and np.array_equal(gd[ko], g["dist"][kr])),
                "contact_sets_equal": bool(sets_equal and np.array_equal(gd[ko] == 0.0, g["dist"][kr] == 0.0))}
    del dc, res
    torch.cuda.empty_cache()
    return out


def block_libccd(args, torch, dist, rank, world, dev):
    """the libccd-style boolean GJK as a batched cross-check of the Jolt kernel."""
    from distance3d_b200 import gjk, random as d3random
    n = 1 << 21
    dc = d3random.random_collider_set_device(args.seed + 23 + 1000 * rank, 2 * n, names=d3random.PRIMITIVES,
                                             center_scale=0.8, device=dev)
    pairs = torch.arange(2 * n, dtype=torch.int32, device=dev).reshape(n, 2)
    res = {}
    ms = timed_steps(torch, dist, world, dev, lambda: res.update(r=gjk.gjk_intersection_libccd_batch(dc, pairs)), 5, 3)
    ms_jolt = timed_steps(torch, dist, world, dev, lambda: res.update(j=gjk.gjk_intersection_batch(dc, pairs)), 5, 3)
    hit, jolt = res["r"][0], res["j"][0]
    out = {"metric": "gjk_intersection_libccd_pairs_per_s", "value": world * n / (ms * 1e-3), "unit": "pairs/s",
           "ms_per_step": ms, "pairs_per_gpu": n, "intersecting_fraction": float(hit.double().mean().item()),
           "jolt_intersection_pairs_per_s": world * n / (ms_jolt * 1e-3),
           "disagreements_with_jolt": int((hit != jolt).sum().item()),
           "includes": "torch.argsort of the (typeA, typeB) keys for the processing order"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cpu_oracle
        m = 200000
        threads = cpu_oracle.max_threads()
        cs = d3random.device_set_to_host(dc, torch.arange(2 * m, device=dev))
        hp = np.arange(2 * m, dtype=np.int32).reshape(m, 2)
        t0 = time.perf_counter()
        ref = cpu_oracle.gjk_intersection_libccd(cs, hp, n_threads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
                               "sample": "first %d pairs of rank 0" % m}
        out["parity_on_cpu_sample"] = {"pairs": m, "booleans_equal": bool(np.array_equal(hit[:m].cpu().numpy(), ref["hit"]))}
    del dc, res
    torch.cuda.empty_cache()
    return out


def block_hydroelastic(args, torch, dist, rank, world, dev):
    """the hydroelastic consumer of the broad phase: boxes of two tetrahedral meshes,
    LBVH over mesh 2 queried with mesh 1, contact plane + polygon for every candidate pair."""
    from distance3d_b200 import hydroelastic_contact as hc
    n = 1 << 20   # tetrahedra per mesh and GPU
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed + 29 + 1000 * rank)
    f64 = dict(device=dev, dtype=torch.float64)
    side = float(n) ** (1.0 / 3.0) * 0.1      # ~ 8 candidate partners per tetrahedron

    def mesh():
        centre = torch.rand((n, 1, 3), generator=gen, **f64) * side
        return (centre + (torch.rand((n, 4, 3), generator=gen, **f64) - 0.5) * 0.1).contiguous(), \
            torch.rand((n, 4), generator=gen, **f64)
    (tp1, e1), (tp2, e2) = mesh(), mesh()
    res = {}
    ms = timed_steps(torch, dist, world, dev, lambda: res.update(r=hc.find_contact_pairs(tp1, e1, tp2, e2)), 5, 3)
    pairs, r = res["r"]
    n_cand = int(pairs.shape[0])
    ms_narrow = timed_steps(torch, dist, world, dev,
                            lambda: hc.intersect_tetrahedron_pairs_batch(pairs, tp1, tp2, e1, e2), 5, 3)
    out = {"metric": "tetrahedron_candidate_pairs_per_s", "value": all_sum(torch, dist, world, dev, n_cand) / (ms * 1e-3),
           "unit": "pairs/s", "ms_per_step": ms, "tetrahedra_per_mesh_per_gpu": n, "candidate_pairs_per_gpu": n_cand,
           "intersecting_fraction": float(r.hit.double().mean().item()),
           "narrow_phase_only_pairs_per_s": all_sum(torch, dist, world, dev, n_cand) / (ms_narrow * 1e-3),
           "stages": "d3d_tetra_aabb x2, d3d_bvh_build, d3d_bvh_overlap (per-thread walk, unsorted queries), "
                     "d3d_tetra_intersect_pairs"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cpu_oracle
        m = min(n_cand, 200000)
        threads = cpu_oracle.max_threads()
        sp = pairs[:m].cpu().numpy()
        u1, i1 = np.unique(sp[:, 0], return_inverse=True)
        u2, i2 = np.unique(sp[:, 1], return_inverse=True)
        hp = np.stack((i1, i2), axis=1).astype(np.int32)
        a1, b1 = tp1[torch.from_numpy(u1).to(dev).long()].cpu().numpy(), e1[torch.from_numpy(u1).to(dev).long()].cpu().numpy()
        a2, b2 = tp2[torch.from_numpy(u2).to(dev).long()].cpu().numpy(), e2[torch.from_numpy(u2).to(dev).long()].cpu().numpy()
        t0 = time.perf_counter()
        ref = cpu_oracle.tetra_pairs(hp, a1, b1, a2, b2, n_threads=threads)
        dt = time.perf_counter() - t0
        g = {k: v[:m].cpu().numpy() for k, v in r.__dict__.items()}
        same = ref["hit"] == g["hit"]
        both = (ref["hit"] == 1) & (g["hit"] == 1)
        out["cpu_baseline"] = {"value": m / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
                               "sample": "first %d candidate pairs of rank 0 (narrow phase only)" % m}
        out["parity_on_cpu_sample"] = {
            "pairs": m, "booleans_equal": bool(same.all()),
            "max_plane_difference": float(np.abs(ref["plane"][both] - g["plane"][both]).max()) if both.any() else 0.0,
            "max_polygon_difference": float(np.abs(ref["polygon"][both] - g["polygon"][both]).max()) if both.any() else 0.0}
    del tp1, tp2, e1, e2, res
    torch.cuda.empty_cache()
    return out


def six_type_mix(args, torch, dist, world, dev):
    """SURVEY 8d "report both": C1 with 10-vertex convex hulls as the sixth type."""
    from distance3d_b200 import gjk, random as d3random
    n = 1 << 20
    dc = d3random.random_collider_set_device(args.seed + 19, 2 * n, names=d3random.PRIMITIVES + ("mesh",),
                                             hull_vertices=(10, 10), device=dev)
    pairs = torch.arange(2 * n, dtype=torch.int32, device=dev).reshape(n, 2)
    out = gjk.gjk_distance_batch(dc, pairs)
    ms = timed_steps(torch, dist, world, dev, lambda: gjk.gjk_distance_batch(dc, pairs, out=out), 5, 3)
    return {"value": world * n / (ms * 1e-3), "unit": "pairs/s", "pairs_per_gpu": n,
            "types": "sphere, ellipsoid, capsule, cylinder, box, 10-vertex convex hull (randn_convex default)",
            "mean_gjk_iterations": float(out.iters.double().mean().item())}


def cpu_baseline(cs, pairs, sample):
    """The C oracle (port of the reference algorithm) on all host threads, bounded sample."""
    from oracle import cpu_oracle
    n = min(sample, len(pairs))
    threads = cpu_oracle.max_threads()
    cpu_oracle.prepare(cs)
    cpu_oracle.gjk_distance(cs, pairs[:min(n, 20000)], n_threads=threads)  # warm
    t0 = time.perf_counter()
    res = cpu_oracle.gjk_distance(cs, pairs[:n], n_threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, threads, n, res


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    n = min(args.cpu_sample, args.pairs)
    cs, pairs = make_workload(args.seed, n)
    from oracle import cpu_oracle
    threads = cpu_oracle.max_threads()
    cpu_oracle.prepare(cs)
    for _ in range(args.warmup):
        cpu_oracle.gjk_distance(cs, pairs[:min(n, 50000)], n_threads=threads)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu_oracle.gjk_distance(cs, pairs, n_threads=threads)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3)
    sample = "%d of %d pairs per step, %d host threads" % (n, args.pairs, threads)
    print(json.dumps({
        "impl": "reference", "metric": "gjk_distance_pairs_per_s", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, pairs_per_gpu):
    return {"workload": "C1 tiled: Jolt GJK distance + closest points, uniform mix of "
                        "{sphere, ellipsoid, capsule, cylinder, box}, default generator scales, "
                        "two fresh colliders per pair (benchmarks/benchmark_gjk.py shape)",
            "pairs_per_gpu": pairs_per_gpu, "seed": args.seed,
            "l2_policy": "inputs (%.0f MB per pass) larger than the 126 MB L2"
                         % (pairs_per_gpu * BYTES_PER_PAIR / 1e6)}


def ncu_traffic_per_pair():
    """DRAM bytes per pair of the dominant GJK kernel from the committed ncu summary
    (profiles/...: `dram__bytes_read.sum`, `dram__bytes_write.sum`, and the `pairs:` line the
    capture script writes).  Returns (bytes per pair | None, source)."""
    path = os.path.join(REPO, NCU_PROFILE)
    try:
        rd = wr = pairs = None
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(path) as f:
            for ln in f:
                parts = ln.split()
                if ln.startswith("pairs:"):
                    pairs = float(parts[1])
                elif parts and parts[0] == "dram__bytes_read.sum" and len(parts) >= 3:
                    rd = float(parts[2].replace(",", "")) * scale.get(parts[1], 1.0)
                elif parts and parts[0] == "dram__bytes_write.sum" and len(parts) >= 3:
                    wr = float(parts[2].replace(",", "")) * scale.get(parts[1], 1.0)
        if rd is None or wr is None or not pairs:
            return None, NCU_PROFILE + " (incomplete)"
        return (rd + wr) / pairs, NCU_PROFILE
    except OSError:
        return None, NCU_PROFILE + " (missing)"


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from distance3d_b200 import _lib, gjk

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    if args.only:
        block = {"broad": lambda: block_broad_phase(args, torch, dist, rank, world, dev, hbm_peak),
                 "epa": lambda: block_epa(args, torch, dist, rank, world, dev),
                 "selfcollision": lambda: block_self_collision(args, torch, dist, rank, world, dev),
                 "pipeline": lambda: block_pipeline(args, torch, dist, rank, world, dev),
                 "libccd": lambda: block_libccd(args, torch, dist, rank, world, dev),
                 "hydroelastic": lambda: block_hydroelastic(args, torch, dist, rank, world, dev)}[args.only]()
        if rank == 0:
            block["n_gpus"] = world
            print(json.dumps(block))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- synthetic shard of this rank --------------------------------------
    cs, pairs = make_workload(args.seed + 1000 * rank, args.pairs)
    dc = cs.device(dev)
    pairs_d = torch.from_numpy(pairs).to(dev)
    n = len(pairs)
    out = None

    def step():
        nonlocal out
        out = gjk.gjk_distance_batch(dc, pairs_d, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        step()                      # keep the GPU under load while nvidia-smi starts
        sampler.wait_samples(2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    n_before = len(sampler.lines)
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    if rank == 0:
        sampler.lines = sampler.lines[n_before:]      # samples taken during the timed region
        if len(sampler.lines) < 2:                    # a very short region: sample the same load again
            for _ in range(args.steps):
                step()
            sampler.wait_samples(2, timeout=2.0)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # mean iterations (must equal the oracle's on the same inputs, checked below)
    res = out.cpu()
    mean_iters = float(res["iters"].mean())
    hit_frac = float((res["dist"] == 0.0).mean())

    # ---- opt-in fp32 arithmetic mode (extra, not the headline) ---------------
    out32 = gjk.gjk_distance_batch(dc, pairs_d, dtype="f32")
    ms32 = timed_steps(torch, dist, world, dev,
                       lambda: gjk.gjk_distance_batch(dc, pairs_d, out=out32, dtype="f32"), 5, 3)
    d32 = out32.dist.cpu().numpy()
    v32 = (out32.status.cpu().numpy() <= 1) & (res["status"] <= 1)
    err32 = np.abs(d32[v32] - res["dist"][v32])
    fp32_mode = {"value": world * n / (ms32 * 1e-3), "unit": "pairs/s", "dtype": "f32",
                 "mean_gjk_iterations": float(out32.iters.double().mean().item()),
                 "abs_distance_error_vs_f64": {"p50": float(np.quantile(err32, 0.5)),
                                               "p999": float(np.quantile(err32, 0.999)),
                                               "max": float(err32.max())}}
    del out32

    # ---- end to end through the public API with HOST buffers ---------------
    # distance3d_b200.stream.GjkDistanceStream: per step the batch's collider records and pairs
    # travel from pinned host memory to the device (compact wire format: a sphere is 4 doubles,
    # not a 4x4 pose), d3d_unpack_colliders + d3d_prepare + d3d_gjk_distance run, and dist /
    # closest points / status travel back; three slots keep PCIe and the SMs busy at once.
    from distance3d_b200 import stream as d3stream
    host = d3stream.pin_batch(cs, pairs, wire=True)
    pipe = d3stream.GjkDistanceStream(len(cs), n, cs.n_vertices, slots=3, device=dev)
    e2e_steps = max(6, min(args.steps, 9))
    staging = [host]

    def e2e_run(k_steps, repack=False):
        pending = []
        last = None
        if repack:
            # The caller's structure-of-arrays set is packed and staged again for every step: a
            # worker thread packs batch k+1 (C++, the GIL is released) while batch k is on its way;
            # four pinned staging sets, so a set is rewritten only after its batch was collected.
            from concurrent.futures import ThreadPoolExecutor
            while len(staging) < 4:
                staging.append(d3stream.pin_batch(cs, pairs, wire=True))
            pool = ThreadPoolExecutor(1)
            fut = pool.submit(d3stream.pin_batch, cs, pairs, True, staging[0])
        for s_ in range(k_steps):
            batch = host
            if repack:
                batch = fut.result()
                if s_ + 1 < k_steps:
                    fut = pool.submit(d3stream.pin_batch, cs, pairs, True, staging[(s_ + 1) % 4])
            pending.append(pipe.submit(batch))
            if len(pending) == len(pipe.slots):
                last = pipe.result(pending.pop(0))
        while pending:
            last = pipe.result(pending.pop(0))
        if repack:
            pool.shutdown()
        return last

    def e2e_timed(k_steps, repack=False):
        barrier()
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(cur)
        r = e2e_run(k_steps, repack)
        pipe.drain_into(cur)
        e1.record(cur)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        # with host-side packing in the loop the device clock does not see the host time: take the
        # larger of the two clocks
        ms = max(e0.elapsed_time(e1), wall if repack else 0.0) / k_steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())

    e2e_res = e2e_run(3)
    e2e_res, e2e_ms = e2e_timed(e2e_steps)
    e2e_value = world * n / (e2e_ms * 1e-3)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_ok = bool(np.array_equal(e2e_res["dist"].numpy(), res["dist"]))
    e2e_run(2, repack=True)
    _, e2e_pack_ms = e2e_timed(6, repack=True)
    del staging[1:]
    # PCIe reference: one large pinned host -> device copy on this GPU
    probe_h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
    probe_d.copy_(probe_h, non_blocking=True)
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(4):
        probe_d.copy_(probe_h, non_blocking=True)
    pe1.record()
    barrier()
    pcie_gbs = 4 * (1 << 28) / (pe0.elapsed_time(pe1) * 1e-3) / 1e9
    del probe_h, probe_d
    del pipe

    # ---- BASELINE configs[1..4] at their stated sizes, on all ranks ----------
    extras = {}
    if not args.no_extra:
        del dc, out, pairs_d, host
        torch.cuda.empty_cache()
        extras["six_type_mix"] = six_type_mix(args, torch, dist, world, dev)
        extras["broad_phase"] = block_broad_phase(args, torch, dist, rank, world, dev, hbm_peak)
        extras["epa"] = block_epa(args, torch, dist, rank, world, dev)
        extras["self_collision"] = block_self_collision(args, torch, dist, rank, world, dev)
        extras["pipeline"] = block_pipeline(args, torch, dist, rank, world, dev)
        extras["libccd"] = block_libccd(args, torch, dist, rank, world, dev)
        extras["hydroelastic"] = block_hydroelastic(args, torch, dist, rank, world, dev)

    if rank == 0:
        fp64_peak = measure_fp64_peak(torch, _lib)
        per_gpu = value / world
        achieved = per_gpu * mean_iters * FLOP_PER_ITER / 1e12
        traffic_per_pair, traffic_src = ncu_traffic_per_pair()
        line = {
            "metric": "gjk_distance_pairs_per_s", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n),
            "mean_gjk_iterations": mean_iters, "intersecting_fraction": hit_frac,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if fp64_peak else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, read from the
                # committed ncu --set full summary (per pair of the captured launch, scaled to this
                # launch); algorithmic bytes are 496 B per pair
                "traffic": None if traffic_per_pair is None else n * traffic_per_pair,
                "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "kernel": "k_gjk_thread<0, primitive instance>",
                "note": "algorithmic flop = pairs x mean_iters x 350 (SURVEY 8d); peak = FP64 FMA "
                        "microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_achieved_gbs": per_gpu * BYTES_PER_PAIR / 1e9,
                "hbm_peak_gbs": hbm_peak,
                "hbm_peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                "hbm_frac": per_gpu * BYTES_PER_PAIR / 1e9 / hbm_peak,
            },
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h),
                    "api": "stream.GjkDistanceStream (3 slots, compact wire records)",
                    "result_equals_device_run": e2e_ok,
                    "h2d_bytes_per_pair": h2d / n, "h2d_bytes_per_pair_reference_layout": 336,
                    "h2d_gbs_per_gpu": h2d / (e2e_ms * 1e-3) / 1e9,
                    "pcie_h2d_peak_gbs_measured": pcie_gbs,
                    "value_including_host_packing": world * n / (e2e_pack_ms * 1e-3),
                    "host_packing": "d3d_pack_wire_host into pinned staging buffers every step (C++, all "
                                    "host threads of the rank); batch k+1 is packed while batch k travels"},
            # k_pair_keys, k_bin_scan, k_bin_scatter, k_gjk_thread x3 (primitive / primitive + hull /
            # all-types instance; an instance whose range is empty exits at once), k_gjk_finish, k_gjk_warp
            "gpu_launches": 8 * args.steps,
            "fp32_mode": fp32_mode,
            "clocks": clocks,
        }
        line.update(extras)
        line["numba_reference"] = NUMBA_REFERENCE
        if not args.no_cpu_baseline:
            cpu_value, threads, sample_n, ref = cpu_baseline(cs, pairs, args.cpu_sample)
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": "pairs/s", "cores": threads, "kind": "port",
                "sample": "first %d of %d pairs of rank 0's shard, C oracle with OpenMP" % (sample_n, n)}
            # the timed output must be the right answer: compare with the oracle sample
            m = ref["status"] <= 1
            line["parity_on_cpu_sample"] = {
                "pairs": int(sample_n),
                "bit_exact_dist": bool(np.array_equal(res["dist"][:sample_n], ref["dist"])),
                "bit_exact_points": bool(np.array_equal(res["closest_a"][:sample_n][m], ref["a"][m])
                                         and np.array_equal(res["closest_b"][:sample_n][m], ref["b"][m])),
                "iters_equal": bool(np.array_equal(res["iters"][:sample_n], ref["iters"])),
            }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
