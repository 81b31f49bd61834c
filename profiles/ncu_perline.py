#!/usr/bin/env python
"""Instructions per unit (e.g. per queue pop) for every source line, in file order.
usage: python profiles/ncu_perline.py prof.ncu-rep <units> [min_inst_per_unit]"""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = ""
tot = 0.0
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or not r[0].strip().isdigit():
        continue
    try:
        i = int(r[7]); s = int(r[4])
    except ValueError:
        continue
    tot += i / units
    if i / units >= thr:
        print("%7.1f inst/unit %6d smp  %s:%-4s %s" % (i / units, s, cur, r[0], r[1].strip()[:105]))
print("total inst/unit %.1f" % tot)
