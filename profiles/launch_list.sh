#!/bin/bash
# launch list of a short C3 bench run (per-launch device times, cold-cache and serialised: compare shares)
tag=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches_bench_c3.csv")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or r[0] == "ID": continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki][:90]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open("gpurun_out/${tag}_launch_shares.txt", "w") as f:
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        line = "%-92s n=%4d total %12.3f ms  share %5.1f %%" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot)
        print(line); f.write(line + "\n")
PY
