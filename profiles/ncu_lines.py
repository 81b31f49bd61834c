#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from an .ncu-rep (cuda,sass correlated view).
usage: python profiles/ncu_lines.py prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
lines = []
cur_file = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 8 or not r[0].strip().isdigit():
        continue
    try:
        lines.append((int(r[4]), int(r[7]), cur_file, int(r[0]), r[1].strip()[:120]))
    except ValueError:
        pass
ts = sum(x[0] for x in lines) or 1
ti = sum(x[1] for x in lines) or 1
print("total stall samples %d, total warp instructions %d" % (ts, ti))
print("--- by instructions executed")
for s, i, f, ln, src in sorted(lines, key=lambda x: -x[1])[:topn]:
    print("%5.1f%% inst %5.1f%% smp  %s:%-4d %s" % (100.0 * i / ti, 100.0 * s / ts, f, ln, src))
