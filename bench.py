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

FLOP_PER_ITER = 350.0  # algorithmic flop per GJK iteration (primitives)
BYTES_PER_PAIR = 496.0  # 336 B in + 160 B out
NCU_DRAM_BYTES_PER_PAIR = (497.15e6 + 228.49e6) / 1048576  # measured, see roofline.traffic


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
    ap.add_argument("--workload", default="gjk", choices=["gjk", "epa", "selfcollision", "pipeline"],
                    help="gjk = headline (C1 + C2 extras); epa = C3; selfcollision = C4; pipeline = C5")
    ap.add_argument("--items", type=int, default=0, help="size of the secondary workloads (0 = default)")
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
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

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


def bench_broad_phase(args, torch, _lib, hbm_peak, steps=5, cpu=True):
    """C2: AABBs of n capsules, LBVH build, all-overlap self query (dense and sparse sets)."""
    from distance3d_b200 import aabb_tree
    n = args.capsules
    out = {"capsules": n}
    for name, scale in (("dense", 2.0), ("constant_density", 2.0 * (n / 2000.0) ** (1.0 / 3.0))):
        cs = make_capsules(n, scale)
        dc = cs.device()
        aabb = _lib.aabb_device(dc)
        bvh = aabb_tree.Lbvh(aabb)
        modes = {}
        for packet in (False, True):   # per-thread vs warp-packet traversal: keep the faster
            tm = []
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pairs, count = bvh.overlap_self(count_visits=True, packet=packet)
                e1.record()
                torch.cuda.synchronize()
                tm.append(e0.elapsed_time(e1))
            modes[packet] = (min(tm), bvh.visits())
            del pairs
        buf = torch.empty((max(count, 1), 2), dtype=torch.int32, device=aabb.device)
        single = {}
        for packet in (False, True):   # single pass: independent traversals vs 8-wide packets
            tm = []
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                bvh.overlap_async(bvh.aabbs, buf, order=bvh.leaf_order(), packet=packet)
                e1.record()
                torch.cuda.synchronize()
                tm.append(e0.elapsed_time(e1))
            bvh.overlap_self(count_visits=True, packet=packet, ordered=False, out=buf)
            single[packet] = (min(tm), bvh.visits())
        packet = single[True][0] < single[False][0]
        visits = single[packet][1]
        tb, tq = [], []
        for it in range(steps + 2):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            bvh.rebuild()
            e[1].record()
            _, cnt_dev = bvh.overlap_async(bvh.aabbs, buf, order=bvh.leaf_order(), packet=packet)  # no host sync
            e[2].record()
            torch.cuda.synchronize()
            assert int(cnt_dev.item()) == count, "single-pass and two-pass overlap counts differ"
            if it >= 2:
                tb.append(e[0].elapsed_time(e[1])); tq.append(e[1].elapsed_time(e[2]))
        tb_ms, tq_ms = float(np.mean(tb)), float(np.mean(tq))
        q_bytes = n * 48.0 + count * 8.0 + visits * 64.0
        # compulsory DRAM traffic: every node record at most once (re-visits hit L1 / L2)
        q_bytes_min = n * 48.0 + count * 8.0 + min(visits, 2 * n - 1) * 64.0
        b_bytes = n * 176.0
        out[name] = {
            "center_scale": scale, "overlap_pairs": int(count), "node_visits": int(visits),
            "traversal": "8-wide packets" if packet else "per thread",
            "single_pass_ms_per_thread_mode": single[False][0], "single_pass_ms_packet_mode": single[True][0],
            "query": "d3d_bvh_overlap: one traversal, warp-staged append (unordered pairs)",
            "ordered_two_pass_ms_per_thread_mode": modes[False][0],
            "ordered_two_pass_ms_packet_mode": modes[True][0],
            "build_ms": tb_ms, "query_ms": tq_ms,
            "build_aabbs_per_s": n / (tb_ms * 1e-3), "overlap_pairs_per_s": count / (tq_ms * 1e-3),
            "queries_per_s": n / (tq_ms * 1e-3),
            "roofline_query": {"bound": "hbm", "achieved": q_bytes / (tq_ms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s",
                               "frac": q_bytes / (tq_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes": "Q*48 + pairs*8 + node_visits*64 (SURVEY 8d); node re-visits "
                                        "are served by L1/L2, so this model can exceed the HBM peak",
                               "frac_compulsory": q_bytes_min / (tq_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes_compulsory": "Q*48 + pairs*8 + min(node_visits, 2n-1)*64"},
            "roofline_build": {"bound": "hbm", "achieved": b_bytes / (tb_ms * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s",
                               "frac": b_bytes / (tb_ms * 1e-3) / 1e9 / hbm_peak,
                               "bytes": "176 B per primitive (SURVEY 8d)"},
        }
        del buf, bvh
        torch.cuda.empty_cache()
        if cpu and name == "constant_density":
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


def secondary_workload(args, torch, dist, rank, world, dev):
    """C3 / C4 / C5 of BASELINE.json (own JSON line, same schema; not the headline)."""
    from distance3d_b200 import _lib, gjk, epa, random as d3random, pipeline
    rs = np.random.RandomState(args.seed + 1000 * rank)
    line = {"n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}
    if args.workload == "epa":
        # C3: intersecting convex hulls with 64-256 vertices, library of shapes, unique poses
        n_pairs = args.items or 500000
        cs = d3random.random_collider_set(rs, 2 * n_pairs, names=("mesh",), center_scale=0.7,
                                          hull_vertices=(64, 256), hull_library=4096)
        pairs = np.arange(2 * n_pairs, dtype=np.int32).reshape(n_pairs, 2)
        dc = cs.device(dev)
        g = gjk.gjk_distance_batch(dc, pairs)
        sel = torch.nonzero((g.dist == 0.0) & (g.n_points == 4)).flatten()
        pairs_d = torch.from_numpy(pairs).to(dev)[sel].contiguous()
        Y = g.simplex[sel].contiguous()
        n = int(sel.numel())
        res = {}
        ms = timed_steps(torch, dist, world, dev, lambda: res.update(r=epa.epa_batch(dc, pairs_d, Y)),
                         args.steps, args.warmup)
        r = res["r"].cpu()
        ms_gjk = timed_steps(torch, dist, world, dev, lambda: gjk.gjk_distance_batch(dc, pairs), 3, 1)
        line.update({"metric": "epa_pairs_per_s", "value": world * n / (ms * 1e-3), "unit": "pairs/s",
                     "ms_per_step": ms,
                     "config": {"workload": "C3: EPA on intersecting convex hulls, 64-256 vertices, "
                                            "4-point GJK simplex", "pairs_per_gpu": n,
                                "hull_pairs_generated": n_pairs},
                     "mean_epa_iterations": float(r["iters"].mean()),
                     "max_faces_assert_rate": float((r["status"] == 7).mean()),
                     "converged_rate": float(r["success"].mean()),
                     "gjk_hull_pairs_per_s": world * n_pairs / (ms_gjk * 1e-3)})
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import cpu_oracle
            m = min(n, 20000)
            threads = cpu_oracle.max_threads()
            cpu_oracle.prepare(cs)
            t0 = time.perf_counter()
            ref = cpu_oracle.epa(cs, pairs_d[:m].cpu().numpy(), Y[:m].cpu().numpy(), n_threads=threads)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": m / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "sample": "first %d EPA pairs" % m}
            ok = ref["status"] != 7
            line["parity_on_cpu_sample"] = {
                "pairs": m, "bit_exact_mtv": bool(np.array_equal(r["mtv"][:m][ok], ref["mtv"][ok])),
                "status_equal": bool(np.array_equal(r["status"][:m], ref["status"]))}
    elif args.workload == "selfcollision":
        # C4: robot arm (6 revolute joints, 8 cylinders), q ~ U(-pi, pi)^6
        from distance3d_b200 import broad_phase, self_collision
        from distance3d_b200.urdf import UrdfTransformManager
        data = os.path.join(REPO, "tests", "data")
        tm = UrdfTransformManager()
        with open(os.path.join(data, "robot_arm.urdf")) as f:
            tm.load_urdf(f.read(), mesh_path=data)
        bvh = broad_phase.BoundingVolumeHierarchy(tm, "robot_arm")
        bvh.fill_tree_with_colliders(tm, fill_self_collision_whitelists=True)
        model = self_collision.RobotModel(tm, bvh)
        n = args.items or 2000000
        q = torch.from_numpy(rs.uniform(-np.pi, np.pi, size=(n, 6))).to(dev)
        res = {}
        ms = timed_steps(torch, dist, world, dev, lambda: res.update(r=model.detect_batch(q)),
                         args.steps, args.warmup)
        mask, n_cand = res["r"]
        line.update({"metric": "self_collision_configurations_per_s", "value": world * n / (ms * 1e-3),
                     "unit": "configurations/s", "ms_per_step": ms,
                     "config": {"workload": "C4: URDF arm (6 joints, 8 cylinders, 17 candidate pairs), "
                                            "FK + AABB + white-list filter + GJK intersection",
                                "configurations_per_gpu": n},
                     "colliding_fraction": float((mask.sum(dim=1) > 0).double().mean().item()),
                     "narrow_phase_candidates_per_configuration": n_cand / n})
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import cpu_oracle
            m = min(n, 200000)
            threads = cpu_oracle.max_threads()
            kin = tm.compile_kinematics(model.frames, "origin")
            qs = q[:m].cpu().numpy()
            t0 = time.perf_counter()
            ref_mask, _ = cpu_oracle.self_collision_masks(model.template, kin, model.pattern, qs, threads)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": m / dt, "unit": "configurations/s", "cores": threads,
                                    "kind": "port", "sample": "first %d configurations" % m}
            line["parity_on_cpu_sample"] = {
                "configurations": m,
                "mask_agreement": float((mask[:m].cpu().numpy() == ref_mask).mean())}
    else:
        # C5: mixed shapes, LBVH broad phase + GJK + EPA
        n = args.items or 2000000
        scale = 0.33 * n ** (1.0 / 3.0)   # ~ 10-30 AABB overlaps per shape
        cs = d3random.random_collider_set(rs, n, names=d3random.PRIMITIVES + ("mesh",),
                                          center_scale=scale, hull_vertices=(10, 10))
        dc = cs.device(dev)
        res = {}
        ms = timed_steps(torch, dist, world, dev,
                         lambda: res.update(r=pipeline.collide(dc, shard=False)), args.steps, args.warmup)
        r = res["r"]
        line.update({"metric": "pipeline_shapes_per_s", "value": world * n / (ms * 1e-3), "unit": "shapes/s",
                     "ms_per_step": ms,
                     "config": {"workload": "C5: mixed random shapes (5 primitives + 10-vertex hulls), LBVH build + "
                                            "all-overlap + GJK distance on candidates + EPA on hits",
                                "shapes_per_gpu": n, "center_scale": scale},
                     "aabb_overlaps_per_shape": r.n_overlaps / n,
                     "candidate_pairs": int(r.candidates.shape[0]),
                     "candidate_pairs_per_s": world * int(r.candidates.shape[0]) / (ms * 1e-3),
                     "contacts": int(r.hits.numel()),
                     "epa_pairs": 0 if r.epa is None else int(r.epa_index.numel())})
        if rank == 0 and not args.no_cpu_baseline:
            # reference algorithms on a bounded sample: incremental tree + stack queries, GJK, EPA
            from oracle import cpu_oracle
            m = min(n, 50000)
            threads = cpu_oracle.max_threads()
            sub = d3random.random_collider_set(np.random.RandomState(args.seed), m,
                                               names=d3random.PRIMITIVES + ("mesh",),
                                               center_scale=0.33 * m ** (1.0 / 3.0), hull_vertices=(10, 10))
            t0 = time.perf_counter()
            A = cpu_oracle.aabb(sub)
            tree = cpu_oracle.Tree()
            tree.insert_aabbs(A)
            pr = tree.query(A)
            cand = pr[pr[:, 0] < pr[:, 1]]
            g = cpu_oracle.gjk_distance(sub, cand, n_threads=threads)
            sel = (g["dist"] == 0.0) & (g["n_points"] == 4)
            cpu_oracle.epa(sub, cand[sel], g["Y"][sel], n_threads=threads)
            dt = time.perf_counter() - t0
            gpu = pipeline.collide(sub, shard=False)
            line["cpu_baseline"] = {"value": m / dt, "unit": "shapes/s", "cores": threads, "kind": "port",
                                    "sample": "%d shapes at the same density (tree build / query single-threaded, "
                                              "GJK / EPA on all threads)" % m}
            line["parity_on_cpu_sample"] = {
                "shapes": m, "candidates_equal": bool(len(cand) == gpu.candidates.shape[0]),
                "contacts_equal": bool(int((g["dist"] == 0.0).sum()) == gpu.hits.numel())}
    if rank == 0:
        print(json.dumps(line))


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
    if args.workload != "gjk":
        secondary_workload(args, torch, dist, rank, world, dev)
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
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop()
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
    # distance3d_b200.stream.GjkDistanceStream: per step the batch's collider arrays and pairs
    # travel from pinned host memory to the device, d3d_prepare + d3d_gjk_distance run, and
    # dist / closest points / status travel back; two slots keep PCIe and the SMs busy at once.
    from distance3d_b200 import stream as d3stream
    host = d3stream.pin_batch(cs, pairs)
    pipe = d3stream.GjkDistanceStream(len(cs), n, cs.n_vertices, slots=2, device=dev)
    e2e_steps = max(4, min(args.steps, 8))

    def e2e_run(k_steps):
        prev = None
        last = None
        for _ in range(k_steps):
            ticket = pipe.submit(host)
            if prev is not None:
                last = pipe.result(prev)
            prev = ticket
        return pipe.result(prev)

    e2e_res = e2e_run(2)
    barrier()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    e2e_res = e2e_run(e2e_steps)
    pipe.drain_into(cur)
    e1.record(cur)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n / (float(t.item()) * 1e-3)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_ok = bool(np.array_equal(e2e_res["dist"].numpy(), res["dist"]))
    del pipe

    if rank == 0:
        fp64_peak = measure_fp64_peak(torch, _lib)
        per_gpu = value / world
        achieved = per_gpu * mean_iters * FLOP_PER_ITER / 1e12
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
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
                # dram__bytes_read.sum + dram__bytes_write.sum of k_gjk_thread<0> from the ncu --set
                # full capture of a 1 Mi-pair launch (profiles/r01_ncu_k_gjk_thread_v6_primitive_instance.txt:
                # 497.1 MB + 228.5 MB incl. the parked simplices), scaled to this launch; algorithmic
                # bytes are 496 B per pair
                "traffic": n * NCU_DRAM_BYTES_PER_PAIR, "traffic_unit": "bytes per launch",
                "kernel": "k_gjk_thread<0, primitive instance>",
                "note": "algorithmic flop = pairs x mean_iters x 350 (SURVEY 8d); peak = FP64 FMA "
                        "microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "hbm_achieved_gbs": per_gpu * BYTES_PER_PAIR / 1e9,
                "hbm_peak_gbs": hbm_peak,
                "hbm_peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                "hbm_frac": per_gpu * BYTES_PER_PAIR / 1e9 / hbm_peak,
            },
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": "stream.GjkDistanceStream (2 slots)",
                    "result_equals_device_run": e2e_ok},
            # k_pair_keys, k_bin_scan, k_bin_scatter, k_gjk_thread x2 (primitive / generic instance),
            # k_gjk_finish, k_gjk_warp
            "gpu_launches": 7 * args.steps,
            "fp32_mode": fp32_mode,
            "clocks": clocks,
        }
        if not args.no_extra:
            del dc, out, pairs_d, host
            torch.cuda.empty_cache()
            line["broad_phase"] = bench_broad_phase(args, torch, _lib, hbm_peak,
                                                    cpu=not args.no_cpu_baseline)
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
