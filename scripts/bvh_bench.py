"""Broad-phase-only timing (C2): python scripts/bvh_bench.py [n]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distance3d_b200 import _lib

class A: pass
args = A(); args.capsules = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
out = bench.bench_broad_phase(args, torch, _lib, 6553.6, steps=5, cpu=False)
for k in ("dense", "constant_density"):
    x = out[k]
    print(k, "build %.3f ms  query %.3f ms  pairs/s %.3e  queries/s %.3e  ordered packet %.2f thread %.2f ms [%s]" % (
        x["build_ms"], x["query_ms"], x["overlap_pairs_per_s"], x["queries_per_s"],
        x["ordered_two_pass_ms_packet_mode"], x["ordered_two_pass_ms_per_thread_mode"], x["traversal"]))
