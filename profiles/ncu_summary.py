#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + the hottest source lines (needs -lineinfo and --import-source on).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for i, h in enumerate(hdr):
    if h in want:
        try:
            v = float(vals[i].replace(",", ""))
            if "stalled" in h and v < 0.05:
                continue
        except ValueError:
            pass
        print("%-80s %-14s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    ci, cs, csrc, cln = col("# Instructions Executed") or col("Instructions Executed"), col("Warp Stall Sampling (All Samples)") or col("# Samples"), col("Source"), col("#")
    print("columns:", h[:12])
    data = []
    for r in rows[1:]:
        try:
            data.append((int(r[cs].replace(",", "")) if cs is not None and r[cs] else 0, int(r[ci].replace(",", "")) if ci is not None and r[ci] else 0, r[cln] if cln is not None else "", r[csrc][:110] if csrc is not None else ""))
        except Exception:
            pass
    tot_s = sum(d[0] for d in data) or 1
    tot_i = sum(d[1] for d in data) or 1
    print("total samples %d, total inst %d" % (tot_s, tot_i))
    for d in sorted(data, reverse=True)[:topn]:
        print("%5.1f%% smp %5.1f%% inst  L%-5s %s" % (100.0 * d[0] / tot_s, 100.0 * d[1] / tot_i, d[2], d[3]))
