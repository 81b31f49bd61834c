#!/usr/bin/env python
"""Main-loop phase split per block (needs a build with -DHP_DBG_MAIN_SPLIT, see profiles/variant.sh):
python profiles/main_split.py [n_blocks] [c2|c3]"""
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
tot = d[:, 0] + d[:, 1]
names = ["real-pop", "score", "vectors->next column", "records", "push+qmin", "prune bookkeeping"]
cols = [8, 9, 5, 6, 4, 10]
print("main loop cycles, all blocks: " + "  ".join("%s %.3g" % (n, d[:, c].sum()) for n, c in zip(names, cols)))
for k in np.argsort(tot)[-5:]:
    exp = d[k, 3] - d[k, 14]
    print("block %4d main %.3g cycles, pops %d (real %d, pruned %d), expansions %d: " % (k, d[k, 1], d[k, 3], d[k, 13], d[k, 14], exp) +
          "  ".join("%s %.0f/exp" % (n, d[k, c] / max(exp, 1)) for n, c in zip(names[1:], cols[1:])) + "  real-pop %.0f/pop (dead discards %.0f each, live pops %.0f each)" % (d[k, 8] / max(d[k, 13], 1), d[k, 15] / max(d[k, 14], 1), (d[k, 8] - d[k, 15]) / max(d[k, 13] - d[k, 14], 1)))
