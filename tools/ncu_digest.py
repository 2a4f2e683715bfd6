#!/usr/bin/env python3
"""Digest an ncu report (made on the GPU box with `ncu --set full`) into the tracked evidence under profiles/:
  python tools/ncu_digest.py profiles/r01 gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]
writes <prefix>_ncu_summary.csv (one row per captured launch) and <prefix>_ncu_traffic.json (per kernel: mean
duration and DRAM bytes per launch), which bench.py reads to fill roofline.traffic."""
import collections
import csv
import json
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    hdr = units = None
    rows = []
    for rep in reps:                      # several captures of the same program (different -k filters) are concatenated
        raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
        part = list(csv.reader(raw.splitlines()))
        if hdr is None:
            hdr, units = part[0], part[1]
            rows = part[:2]
        col = {h: i for i, h in enumerate(part[0])}
        for r in part[2:]:
            rows.append([r[col[h]] if h in col else "" for h in hdr])
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(prefix + "_ncu_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    unit = {h: u for h, u in zip(hdr, units)}
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].strip()
        g = r[hdr.index("Grid Size")]
        dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * scale.get(unit["gpu__time_duration.sum"], 1.0)
        rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) * scale.get(unit["dram__bytes_read.sum"], 1.0)
        wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) * scale.get(unit["dram__bytes_write.sum"], 1.0)
        ins = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", "") or 0)
        issue = float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")].replace(",", "") or 0)
        agg[name].append({"grid": g, "us": dur, "dram_read": rd, "dram_write": wr, "inst": ins, "issue": issue})
    out = {}
    for k, v in agg.items():
        def vol(x):
            d = [int(t) for t in x["grid"].strip("() ").split(",")]
            return d[0] * d[1] * d[2]
        big = [x for x in v if vol(x) == max(vol(y) for y in v)]            # the batch-sized launches
        out[k] = {"launches": len(big), "grid": big[0]["grid"], "mean_us": sum(x["us"] for x in big) / len(big),
                  "dram_bytes_per_launch": sum(x["dram_read"] + x["dram_write"] for x in big) / len(big),
                  "warp_inst_per_launch": sum(x["inst"] for x in big) / len(big),
                  "issue_active_pct": sum(x["issue"] for x in big) / len(big)}
    # stamp with the digest of the kernel sources the capture was made from: bench.py ignores a stale file
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out["csrc_sha"] = bench.csrc_sha()
    out["file_sha"] = bench.csrc_file_sha()          # per file: a kernel's figures stay valid while its own sources do
    json.dump(out, open(prefix + "_ncu_traffic.json", "w"), indent=1)
    del out["csrc_sha"], out["file_sha"]
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["mean_us"]):
        print("%-24s %8.1f us  dram %8.2f MB  grid %s" % (k, v["mean_us"], v["dram_bytes_per_launch"] / 1e6, v["grid"]))


if __name__ == "__main__":
    main()
