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
# round 2: streaming lanes + shared slab pool, bulk-async prep on the C3 stream (incl. a block whose cells end the arrays), the
# realignment pipeline, the CIGAR projection, 4-word node masks, the block service
from hiphase_b200 import _abi as A
ctx.set_lanes(3)
bs = [synth.config_c3_stream(10, first_block=10 * i) for i in range(4)]
hs = [ctx.astar_submit(x) for x in bs[:3]]
for h, x in zip(hs, bs[:3]):
    o = ctx.astar_wait(h); r = O.astar_solve(x, threads=4, want_heuristic=False, want_counters=False)
    assert np.array_equal(o.h1, r.h1) and np.array_equal(o.stats, r.stats)
d, vt = synth.config_realign(1, 0, window=6000, n_var=10, n_hom=4, n_reads=16, read_lo=800, read_hi=2000, sv_max=200, err=0.004)
rb = A.RealignBatch(global_failure_minimum=2, global_failure_ratio=0.2, **d)
c2 = lib.Context(A.hp_params(1000, 3, 500, 6), device=0)
ro = c2.realign_block_batch(rb); rr = O.realign_block_batch(rb, A.hp_params(1000, 3, 500, 6))
assert np.array_equal(ro.map_mode, rr.map_mode) and np.array_equal(ro.alleles[: int(rr.as_struct().assembled.n_cells)], rr.alleles[: int(rr.as_struct().assembled.n_cells)])
c2.close()
import importlib.util
spec = importlib.util.spec_from_file_location("t", os.path.join(R, "tests", "test_gpu_wfa.py")); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
big = t._dense_snv_batch(360, 2, seed=1)
assert (ctx.wfa_align_batch(big, trav_words=32).status == 0).all()
svc = lib.BlockService(device=0)
h1, h2, st = svc.solve_one(bs[3], 2)
svc.close()
print("round-2 sanitizer workload ok")
# second half of round 2: the piece filter (hash set of read pieces, bounded depth-first walk), the warp-wide extension and the
# atomic-free slot insertion, on noisy reads around max_edit_distance
wb2, _, _ = synth.config_c4(1, window=8000, n_het=14, n_hom=14, n_reads=10, read_lo=2200, read_hi=3500, sv_max=200, err=0.01, p_noisy=0.5, err_noisy=0.3)
p2 = A.hp_params(1000, 3, 500, 100)
c3 = lib.Context(p2, device=0)
wo2 = c3.wfa_align_batch(wb2, trav_words=2); wr2 = O.wfa_align(wb2, p2, threads=4, trav_words=2)
assert np.array_equal(wo2.status, wr2.status) and np.array_equal(wo2.score, wr2.score) and np.array_equal(wo2.alleles, wr2.alleles)
assert c3.wfa_filtered() > 0 and (wr2.status == 0).any()
c3.close()
print("round-2 piece-filter workload ok")
