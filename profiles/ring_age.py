"""How old is the expansion that created a plane-rescored pop of the main loop?  (sizes an expansion-vector ring.)
Needs a build with -DHP_DBG_RING_AGE -DHP_DBG_MAIN_SPLIT: profiles/variant.sh ring -DHP_DBG_RING_AGE -DHP_DBG_MAIN_SPLIT;
HP_B200_LIB=build/libhp_ring.so python profiles/ring_age.py"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiphase_b200 import lib, synth
for name, batch in (("c2", synth.config_c2(1000)), ("c3", synth.config_c3(600))):
    nb = batch.n_blocks
    ctx = lib.Context(device=0)
    L = lib.lib()
    L.hp_debug_enable_block_cycles(ctx.handle, 1)
    out = ctx.astar_solve_batch(batch, want_counters=True)
    d = np.zeros(nb * 16, np.uint64)
    assert L.hp_debug_read_block_cycles(ctx.handle, d.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
    d = d.reshape(nb, 16).astype(np.float64)
    n = d[:, 11].sum()
    print(name, "plane-rescored pops %d; parent expansion within 4: %.1f%%  8: %.1f%%  16: %.1f%%  32: %.1f%%" % (n, 100*d[:,8].sum()/n, 100*d[:,9].sum()/n, 100*d[:,10].sum()/n, 100*d[:,12].sum()/n))
    k = np.argsort(d[:, 1])[-3:]
    for b in k: print("   straggler main %.3g: planes %d within 16: %.1f%% 32: %.1f%%" % (d[b,1], d[b,11], 100*d[b,10]/max(d[b,11],1), 100*d[b,12]/max(d[b,11],1)))
    ctx.close()
