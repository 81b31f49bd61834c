#!/usr/bin/env python
"""Row f1 bench: local realignment jobs/s, GPU (hp_local_realign_batch, host buffers, end to end) vs the CPU oracle.
usage: python profiles/bench_local.py [n_blocks] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 40
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
batch = synth.config_local(nb, full_rows=True)
ctx = lib.Context(device=0)
out = ctx.local_realign_batch(batch)
ts, ks = [], []
for _ in range(reps):
    t0 = time.perf_counter(); out = ctx.local_realign_batch(batch); ts.append(time.perf_counter() - t0); ks.append(ctx.last_kernel_ms())
t0 = time.perf_counter(); ref = O.local_realign(batch); tc = time.perf_counter() - t0
same = all(np.array_equal(getattr(out, k), getattr(ref, k)) for k in ("alleles", "quals", "match_class", "edit_distance", "status"))
cells = int(batch.row_off[-1])
print("local realignment: %d jobs, %d cells, %d read bases, %d aligned segments; bit-exact vs oracle: %s" % (batch.n_jobs, cells, len(batch.read_bytes), len(batch.seg_len), same))
print("GPU end to end %.2f ms (%.0f jobs/s), kernel %.3f ms; CPU oracle (1 thread) %.1f ms (%.0f jobs/s)" %
      (1e3 * np.median(ts), batch.n_jobs / np.median(ts), np.median(ks), 1e3 * tc, batch.n_jobs / tc))
