"""Graph-WFA kernel timeline: per-job start / end times (profiling build mode 4) of one C4 chunk launch.
usage: python profiles/wfa_timeline.py [n_blocks]"""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from hiphase_b200 import lib, _abi as A
from profiles import bench_c4
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 250
(d, vt), = bench_c4.generate(0, nb, nb, os.cpu_count())
b = A.RealignBatch(**d)
ctx = lib.Context(device=0)
ctx.set_wfa_filter(os.environ.get("HP_WFA_FILTER", "1") != "0")
ctx.wfa_align_batch(b.wfa)
ctx.wfa_align_batch(b.wfa)
print("production launch: %.2f ms, filter answered %d" % (ctx.last_kernel_ms(), ctx.wfa_filtered()))
ctx.set_wfa_build_mode(4)
o = ctx.wfa_align_batch(b.wfa, want_counters=True)
print("timed launch: %.2f ms" % ctx.last_kernel_ms())
c = o.counters
t0 = c["set_ops"].astype(np.int64); t1 = c["n_nodes"].astype(np.int64)
base = t0.min()
t0 = (t0 - base) / 1e6; t1 = (t1 - base) / 1e6
dur = t1 - t0
maxed = o.status == A.HP_WFA_MAX_EDIT_DISTANCE
print("jobs %d, MaxEditDistance %d, span %.2f ms" % (len(dur), maxed.sum(), t1.max()))
print("MaxED jobs: duration ms percentiles 5/50/95/max:", np.percentile(dur[maxed], [5, 50, 95, 100]).round(2), "waves mean %.0f" % c["waves_processed"][maxed].mean())
print("           starts (ms) 50/95/max:", np.percentile(t0[maxed], [50, 95, 100]).round(2), " ends 50/95/max:", np.percentile(t1[maxed], [50, 95, 100]).round(2))
ok = ~maxed
print("other jobs: duration ms 5/50/95/max:", np.percentile(dur[ok], [5, 50, 95, 100]).round(3), "sum %.1f s of warp time; MaxED sum %.1f s" % (dur[ok].sum() / 1e3, dur[maxed].sum() / 1e3))
edges = np.arange(0, t1.max() + 2, 2.0)
for lo, hi in zip(edges[:-1], edges[1:]):
    m = ok & (t0 >= lo) & (t0 < hi)
    running = ((t0 < hi) & (t1 > lo)).sum()
    print("  started in [%4.0f, %4.0f) ms: %6d other jobs, mean duration %.3f ms, mean waves %.0f; jobs alive in the window %d (MaxED %d)"
          % (lo, hi, m.sum(), dur[m].mean() if m.any() else 0, c["waves_processed"][m].mean() if m.any() else 0, running, (maxed & (t0 < hi) & (t1 > lo)).sum()))
