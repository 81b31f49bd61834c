"""Mixed small workload (A*, graph-WFA, local realignment) checked against the oracle; run under compute-sanitizer:
  compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize_workload.py"""
import sys; import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import oracle_lib as O
from hiphase_b200 import lib, synth
ctx = lib.Context(device=0)
b = synth.config_c2(n_blocks=24, n_var=120, n_reads=40)
o = ctx.astar_solve_batch(b, want_heuristic=True, want_counters=True); r = O.astar_solve(b)
assert np.array_equal(o.h1, r.h1) and np.array_equal(o.stats, r.stats) and np.array_equal(o.counters, r.counters)
o = ctx.astar_solve_batch(b)
b3 = synth.config_c3(12, first_block=40)
o = ctx.astar_solve_batch(b3); r = O.astar_solve(b3, threads=4, want_heuristic=False, want_counters=False)
assert np.array_equal(o.h1, r.h1) and np.array_equal(o.stats, r.stats)
wb, jb, meta = synth.config_c4(1, window=12000, n_het=20, n_hom=20, n_reads=10, read_lo=1500, read_hi=3000, sv_max=300)
wo = ctx.wfa_align_batch(wb, trav_words=4); wr = O.wfa_align(wb, trav_words=4)
assert np.array_equal(wo.alleles, wr.alleles) and np.array_equal(wo.score, wr.score)
lb = synth.config_local(1, full_rows=True, n_reads=12)
lo = ctx.local_realign_batch(lb); lr = O.local_realign(lb)
assert np.array_equal(lo.alleles, lr.alleles) and np.array_equal(lo.quals, lr.quals)
print("sanitizer workload ok")
