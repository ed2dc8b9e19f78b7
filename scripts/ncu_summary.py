#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (key raw metrics + hottest source lines).

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines] > profiles/rNN_<kernel>.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread ", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct",
        "gpu__dram_throughput.avg.pct", "sass__inst_executed_local"]


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name)
    print("report:", rep)
    print("\n== raw metrics ==")
    for h, u, v in zip(hdr, units, vals):
        if any(k in (h + " ") for k in KEYS):
            try:
                if float(v.replace(",", "")) == 0.0:
                    continue
            except ValueError:
                pass
            print("%-90s %-12s %s" % (h, u, v))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    cur, hdr, agg = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0] != "":
            d = dict(zip(hdr, r))
            try:
                agg.append((int(d["# Samples"]), int(d["Instructions Executed"]),
                            int(d["Thread Instructions Executed"]), cur, r[0], r[1].strip()[:90]))
            except ValueError:
                pass
    ts, ti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    by_file = collections.Counter()
    for a in agg:
        by_file[a[3]] += a[0]
    print("\n== stall samples by file ==")
    print(dict(by_file))
    print("\n== hottest source lines (stall samples, warp instructions, active threads per instruction) ==")
    for a in sorted(agg, reverse=True)[:n_lines]:
        print("%5.1f%% samp %5.1f%% inst thr/inst=%5.1f  %s:%s  %s"
              % (a[0] * 100 / ts, a[1] * 100 / ti, a[2] / max(a[1], 1), a[3], a[4], a[5]))


if __name__ == "__main__":
    main()
