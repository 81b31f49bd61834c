"""Sanitizer workload for the last changes of round 2: packed sub-solver vectors (both builds of the solver kernels: the dense one
is forced on a launch alone with HP_DBG_DENSE=2), streaming launches that take the dense build by themselves, and the multi-panel
edit distance (hbuf row in global memory).  Everything is checked against the oracle.
  compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize_r2j.py"""
import sys; import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import oracle_lib as O
from hiphase_b200 import lib, synth
import test_local_realign as T
ctx = lib.Context(device=0)
b3 = synth.config_c3(10, first_block=40)
r = O.astar_solve(b3, threads=4, want_heuristic=False, want_counters=False)
for mode in ("1", "2"):
    os.environ["HP_DBG_DENSE"] = mode
    o = ctx.astar_solve_batch(b3)
    assert np.array_equal(o.h1, r.h1) and np.array_equal(o.h2, r.h2) and np.array_equal(o.stats, r.stats), mode
ctx.set_lanes(3)                          # three small batches in flight, still forced onto the dense build (a natural choice of
ids = np.arange(0, 90, dtype=np.uint64)   # it needs thousands of blocks per launch: too slow under the sanitizer)
parts = [ids[k::3] for k in range(3)]
bs = [synth.stream_blocks(p) for p in parts]
hs = [ctx.astar_submit(x) for x in bs]
outs = [ctx.astar_wait(h) for h in hs]
for p, x, o in zip(parts, bs, outs):
    sel = np.arange(0, len(p), max(1, len(p) // 12))
    sub = x.select(sel)
    rr = O.astar_solve(sub, threads=4, want_heuristic=False, want_counters=False)
    assert np.array_equal(o.stats[sel], rr.stats)
os.environ.pop("HP_DBG_DENSE")
print("solver builds ok")
rng = np.random.default_rng(5)
letters = np.frombuffer(b"ACGT", np.uint8)
a = letters[rng.integers(0, 4, 16500)]
bm = T._mutated(rng, a)
d = ctx.edit_distance_batch([(a, bm), (a[:16384], bm)])
assert int(d[0]) == O.edit_distance(a, bm) and int(d[1]) == O.edit_distance(a[:16384], bm)
ref = letters[rng.integers(0, 4, 240)].tobytes()
ins = letters[rng.integers(0, 4, 16600)]
sv = T.Variant(0, 4, 100, 1, ref[100:101], ref[100:101] + ins.tobytes())
got = T._mutated(rng, ins).tobytes()
read = ref[:101] + got + ref[101:]
job = T._single_job([sv], 0, [(0, 0, 101), (101, 101 + len(got), len(ref) - 101)], read, [30] * len(read))
T._same(ctx.local_realign_batch(job), O.local_realign(job))
print("multi-panel edit distance ok")
ctx.close()
