#!/usr/bin/env python
"""Per-block cycles of the heuristic pre-pass vs the main A* loop (counting kernel variant).
usage: python profiles/block_cycles.py [n_blocks] [c2|c3] [team]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = synth.config_c2(nb) if len(sys.argv) < 3 or sys.argv[2] == "c2" else synth.config_c3(nb)
ctx = lib.Context(device=0)
if len(sys.argv) > 3:
    ctx.set_team(int(sys.argv[3]))
L = lib.lib()
L.hp_debug_enable_block_cycles(ctx.handle, 1)
for _ in range(2):
    out = ctx.astar_solve_batch(batch, want_counters=True)
d = np.zeros(nb * 16, np.uint64)
assert L.hp_debug_read_block_cycles(ctx.handle, d.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
d = d.reshape(nb, 16).astype(np.float64)
nvar = np.diff(batch.var_off.astype(np.int64))
print("kernel ms %.3f  team %d  variants committed per round: mean %.2f" % (ctx.last_kernel_ms(), d[0, 5], (nvar / d[:, 4]).mean()))
print("pre-pass: cycles/pop mean %.0f  cycles/block mean %.3g max %.3g" % ((d[:, 0] / d[:, 2]).mean(), d[:, 0].mean(), d[:, 0].max()))
print("main    : cycles/pop mean %.0f  cycles/block mean %.3g max %.3g" % ((d[:, 1] / np.maximum(d[:, 3], 1)).mean(), d[:, 1].mean(), d[:, 1].max()))
tot = d[:, 0] + d[:, 1]
print("sum pre-pass %.3g  sum main %.3g  slowest block %.3g cycles" % (d[:, 0].sum(), d[:, 1].sum(), tot.max()))
mp = d[:, 8:11].sum(0)
print("main split (all blocks): real-pop %.3g  expand %.3g  rest %.3g cycles | real pops %d  pruned %d  plane-rescored expansions %d (%.3g cycles)"
      % (mp[0], mp[1], mp[2], d[:, 13].sum(), d[:, 14].sum(), d[:, 11].sum(), d[:, 12].sum()))
print("5 slowest blocks: [total, pre-pass, main cycles | pops pre-pass, main | variants/round] then main split "
      "[real-pop, expand, rest (records+push+prune) cycles | real pops, pruned, final queue]")
for k in np.argsort(tot)[-5:]:
    print("  ", [int(tot[k]), int(d[k, 0]), int(d[k, 1])], [int(d[k, 2]), int(d[k, 3])], "%.2f" % (nvar[k] / d[k, 4]),
          d[k, 8:11].astype(np.int64).tolist(), d[k, 13:16].astype(np.int64).tolist(), "planes", d[k, 11:13].astype(np.int64).tolist())
