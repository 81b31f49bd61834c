#!/usr/bin/env python
"""Sub-solver (heuristic pre-pass) phase split of warp 0 per block (needs a build with -DHP_DBG_SUB_SPLIT):
python profiles/sub_split.py [n_blocks] [c2|c3]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = synth.config_c2(nb) if len(sys.argv) < 3 or sys.argv[2] == "c2" else synth.config_c3(nb)
ctx = lib.Context(device=0)
L = lib.lib()
L.hp_debug_enable_block_cycles(ctx.handle, 1)
for _ in range(2):
    out = ctx.astar_solve_batch(batch, want_counters=True)
d = np.zeros(nb * 16, np.uint64)
assert L.hp_debug_read_block_cycles(ctx.handle, d.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
d = d.reshape(nb, 16).astype(np.float64)
nvar = np.diff(batch.var_off.astype(np.int64))
tot = d[:, 0] + d[:, 1]
print("kernel %.3f ms; warp-0 sub-solver cycles, all blocks: pop %.3g  re-seat %.3g  score %.3g  rest %.3g | expansions %.3g real pops %.3g plane-rescored %.3g"
      % (ctx.last_kernel_ms(), d[:, 8].sum(), d[:, 9].sum(), d[:, 10].sum(), d[:, 11].sum(), d[:, 14].sum(), d[:, 12].sum(), d[:, 13].sum()))
for k in np.argsort(tot)[-5:]:
    e, r = max(d[k, 14], 1), max(d[k, 12], 1)
    print("block %5d N %4d pre-pass %.3g cycles, rounds %d: warp 0 expansions %d (real pops %d, plane-rescored %d): pop %.0f/real  re-seat %.0f/real  score %.0f/exp  rest %.0f/exp  | sum %.3g | pop = push-back %.0f + owner/entry %.0f + remove/rescan/min %.0f | warp 0: in sub-solves %.3g, waiting at the round barrier %.3g cycles"
          % (k, nvar[k], d[k, 0], d[k, 4], d[k, 14], d[k, 12], d[k, 13], d[k, 8] / r, d[k, 9] / r, d[k, 10] / e, d[k, 11] / e, d[k, 8:12].sum(),
             d[k, 15] / r, (d[k, 7] - d[k, 15]) / r, (d[k, 8] - d[k, 7]) / r, d[k, 5], d[k, 6]))
