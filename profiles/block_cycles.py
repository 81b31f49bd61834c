#!/usr/bin/env python
"""Per-block cycles of the heuristic pre-pass vs the main A* loop (counting kernel variant), C2 workload."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = synth.config_c2(nb) if len(sys.argv) < 3 or sys.argv[2] == 'c2' else synth.config_c3(nb)
ctx = lib.Context(device=0)
if len(sys.argv) > 3:
    ctx.set_team(int(sys.argv[3]))
L = lib.lib()
L.hp_debug_enable_block_cycles(ctx.handle, 1)
for _ in range(2):
    out = ctx.astar_solve_batch(batch, want_counters=True)
d = np.zeros(nb * 8, np.uint64)
assert L.hp_debug_read_block_cycles(ctx.handle, d.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
d = d.reshape(nb, 8).astype(np.float64)
nvar = np.diff(batch.var_off.astype(np.int64))
print('team', d[0, 6], 'variants/round mean %.2f' % (nvar / d[:, 4]).mean(), 'warp0 barrier wait share %.2f' % (d[:, 5].sum() / d[:, 0].sum()))
print("kernel ms", ctx.last_kernel_ms())
print("sub : cycles/pop mean %.0f  total cycles mean %.3g max %.3g" % ((d[:, 0] / d[:, 2]).mean(), d[:, 0].mean(), d[:, 0].max()))
mp = np.maximum(d[:, 3], 1)
print("main: cycles/pop mean %.0f  total cycles mean %.3g max %.3g" % ((d[:, 1] / mp).mean(), d[:, 1].mean(), d[:, 1].max()))
tot = d[:, 0] + d[:, 1]
i = np.argsort(tot)[-5:]
print("slowest blocks: variants/round", (nvar / d[:, 4])[i], "\n  total cycles", tot[i], "sub", d[i, 0], "main", d[i, 1], "pops sub/main", d[i, 2], d[i, 3])
print("sum sub %.3g sum main %.3g" % (d[:, 0].sum(), d[:, 1].sum()))
