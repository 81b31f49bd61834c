#!/usr/bin/env python
"""Per-block cycles on the C3 stream (first n blocks), per team size: total work vs the slowest chain.
usage: python profiles/stream_cycles.py [n_blocks] [out_prefix|-] [teams, e.g. 1,2]"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiphase_b200 import lib, synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
prefix = sys.argv[2] if len(sys.argv) > 2 else None
batch = synth.config_c3_stream(nb)
nvar = np.diff(batch.var_off.astype(np.int64))
_, noisy = synth.stream_headers(0, nb)
L = lib.lib()
for team in ([int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else (1, 2, 4)):
    ctx = lib.Context(device=0)
    ctx.set_team(team)
    for _ in range(2):
        t0 = time.perf_counter(); ctx.astar_solve_batch(batch); dt = time.perf_counter() - t0
    kms = ctx.last_kernel_ms()
    L.hp_debug_enable_block_cycles(ctx.handle, 1)
    out = ctx.astar_solve_batch(batch, want_counters=True)
    d = np.zeros(nb * 16, np.uint64)
    assert L.hp_debug_read_block_cycles(ctx.handle, d.ctypes.data_as(C.POINTER(C.c_uint64)), nb) == 0
    d = d.reshape(nb, 16).astype(np.float64)
    tot = d[:, 0] + d[:, 1]
    top = np.argsort(tot)[::-1]
    print("team %d: production kernels %.1f ms (e2e %.1f ms) | counting run %.1f ms | sum cycles %.4g (pre-pass %.4g main %.4g) "
          "slowest %.4g | clean blocks sum %.4g, noisy blocks sum %.4g"
          % (team, kms, dt * 1e3, ctx.last_kernel_ms(), tot.sum(), d[:, 0].sum(), d[:, 1].sum(), tot.max(), tot[noisy == 0].sum(), tot[noisy == 1].sum()))
    print("   top blocks [id, N, noisy, Mcycles]:", [(int(k), int(nvar[k]), int(noisy[k]), round(tot[k] / 1e6)) for k in top[:12]])
    print("   blocks over 50 / 100 / 200 Mcycles:", int((tot > 50e6).sum()), int((tot > 100e6).sum()), int((tot > 200e6).sum()))
    if prefix and prefix != '-':
        np.save("%s_team%d.npy" % (prefix, team), d)
    ctx.close()
